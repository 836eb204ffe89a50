"""Diagnostic: two replicas as THREADS of one process on one GPU, one zero-copy exchange; on an arrival
timeout dump the arrival flags of both heaps as seen through both mappings."""
import ctypes as C
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.update(TOK_BARRIER_TIMEOUT_MS="4000", TOK_MAX_CTAS="16", TOK_STAGING_MB="32", TOK_SYMM_POOL_MB="64")

import torch  # noqa: E402

from torch_on_k8s_b200 import _ffi  # noqa: E402
from torch_on_k8s_b200.comm import Communicator  # noqa: E402
from torch_on_k8s_b200.elastic_dp import symm_tensor  # noqa: E402

path = os.path.join(tempfile.mkdtemp(), "r")
comms, bufs, streams = {}, {}, {}
delay = float(os.environ.get("DELAY", "0.5"))
staged_first = int(os.environ.get("STAGED_FIRST", "1"))


def peek(c, rank, off=32768, n=8):
    a = (C.c_uint32 * n)()
    _ffi.check(_ffi.lib().tok_comm_debug_peek(c._h, rank, off, a, n))
    return list(a)


def body(r):
    torch.cuda.set_device(0)
    comms[r] = Communicator("diag", r, 2, 0, rendezvous_path=path)
    streams[r] = torch.cuda.Stream()
    with torch.cuda.stream(streams[r]):
        bufs[r] = symm_tensor(comms[r], 1 << 18, torch.bfloat16).fill_(float(r + 1))
        if staged_first:
            x = torch.ones(200003, device="cuda", dtype=torch.bfloat16)
            comms[r].allreduce_bucket(x, x, scale=0.5, algo=3, stream=streams[r])
        streams[r].synchronize()
        if r == 0:
            time.sleep(delay)
        comms[r].allreduce_bucket(bufs[r], bufs[r], scale=0.5, algo=3, stream=streams[r])
        streams[r].synchronize()


ts = [threading.Thread(target=body, args=(r,)) for r in range(2)]
[t.start() for t in ts]
[t.join(60) for t in ts]
for r in range(2):
    try:
        comms[r].status()
        print("rank", r, "ok", float(bufs[r][0]), comms[r].last_algo())
    except Exception as e:  # noqa: BLE001
        print("rank", r, "ERR", str(e)[:200])
for r in range(2):
    for q in range(2):
        print("rank %d's mapping of rank %d arr flags:" % (r, q), peek(comms[r], q),
              "flag[0][*]:", peek(comms[r], q, 0, 4))
