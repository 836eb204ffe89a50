"""Import shim: the product package lives in ``torch-on-k8s_b200/`` (hyphenated like the reference
repo's name, so not importable by that name).  ``import torch_on_k8s_b200.comm`` resolves there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "torch-on-k8s_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
