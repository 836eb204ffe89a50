"""The plain-C oracle (oracle/allreduce_oracle.c) against the numpy oracle, bit for bit, for every
dtype triple, world 1..8, PRE/POST scale — two independent restatements of the same path.  CPU only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import harness
from oracle import allreduce_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = {"f32": 0, "bf16": 1, "f16": 2}


@pytest.fixture(scope="module")
def clib():
    lib_path = subprocess.check_output([sys.executable, os.path.join(ROOT, "oracle", "build_oracle.py")],
                                       text=True).strip()
    lib = C.CDLL(lib_path)
    lib.oracle_allreduce.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_size_t, C.c_int, C.c_int,
                                     C.c_int, C.c_float, C.c_int, C.c_void_p]
    lib.oracle_allreduce.restype = C.c_int
    lib.oracle_f32_to_f16.argtypes = [C.c_float]
    lib.oracle_f32_to_f16.restype = C.c_uint16
    lib.oracle_f32_to_bf16.argtypes = [C.c_float]
    lib.oracle_f32_to_bf16.restype = C.c_uint16
    return lib


def c_allreduce(lib, inputs, din, dw, dout, scale, post):
    n = len(inputs[0])
    store = {"f32": np.float32, "bf16": np.uint16, "f16": np.float16}
    arrs = [np.ascontiguousarray(a, dtype=store[din]) for a in inputs]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    out = np.zeros(n, dtype=store[dout])
    rc = lib.oracle_allreduce(ptrs, len(arrs), n, CODE[din], CODE[dw], CODE[dout], scale, int(post),
                              out.ctypes.data)
    assert rc == 0
    return out


@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("post", [False, True])
def test_c_oracle_equals_numpy_oracle(clib, world, post):
    seed = 40 + world
    for din in O.DTYPES:
        for dw in O.DTYPES:
            for dout in O.DTYPES:
                for pattern in ("randn", "wide"):
                    seed += 1
                    ins = [harness.gen_input(seed, r, 3001, din, pattern) for r in range(world)]
                    scale = np.float32(1.0 / 3.0) if post else np.float32(1.0 / world)
                    want = O.allreduce_oracle(ins, din, dw, dout, scale, post)
                    got = c_allreduce(clib, ins, din, dw, dout, float(scale), post)
                    assert harness.bits_equal(got, want, dout), (din, dw, dout, pattern)


def test_c_conversions_match_numpy(clib):
    rs = np.random.RandomState(1)
    x = (rs.standard_normal(50000) * np.exp(rs.uniform(-30, 15, 50000))).astype(np.float32)
    edge = np.array([0.0, -0.0, 65504, 65519.99, 65520, 1e-8, 5.96e-8, 2.9802322e-8, 6.1e-5, np.inf,
                     -np.inf], dtype=np.float32)
    x = np.concatenate([x, edge])
    with np.errstate(over="ignore"):
        want16 = x.astype(np.float16).view(np.uint16)
    got16 = np.array([clib.oracle_f32_to_f16(float(v)) for v in x], dtype=np.uint16)
    assert np.array_equal(got16, want16)
    gotb = np.array([clib.oracle_f32_to_bf16(float(v)) for v in x], dtype=np.uint16)
    assert np.array_equal(gotb, O.f32_to_bf16_bits(x))
