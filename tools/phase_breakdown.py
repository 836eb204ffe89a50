"""Per-phase timing of the exchange kernels (TOK_DEBUG_PHASES=1): for each algorithm and size, the
median over CTAs and ranks of {stage, barrier A, reduce, barrier B, gather/copy-out}.  The zero-copy
kernels (rows *_inplace) have no stage / barrier A / copy-out: reduce+push, then the trailing barrier
(barB); the arrival kernel in front of them is a separate launch and is not in `total`.  torchrun, one
rank per GPU.  Not part of the product."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["TOK_DEBUG_PHASES"] = "1"

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from torch_on_k8s_b200 import _ffi  # noqa: E402
from torch_on_k8s_b200.comm import Communicator  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = []
    for ctas in [int(x) for x in os.environ.get("CTAS", "64").split(",")]:
        os.environ["TOK_MAX_CTAS"] = str(ctas)
        comm = Communicator("phase", rank, world, local,
                            rendezvous_path="/tmp/tok8s-phase-%s-%d" % (os.environ["MASTER_PORT"], ctas))
        for mb in [float(x) for x in os.environ.get("SIZES_MB", "4,32,128").split(",")]:
            n = int(mb * (1 << 20)) // 2
            staged = [torch.full((n,), float(rank + 1), dtype=torch.bfloat16, device="cuda") for _ in range(4)]
            pooled = [comm.symm_empty(n, torch.bfloat16).fill_(float(rank + 1)) for _ in range(4)]
            for algo, zc in ((3, False), (4, False), (3, True), (4, True)):
                bufs = pooled if zc else staged
                if algo == 4 and not comm.caps().multicast:
                    continue
                rows = []
                for it in range(8):
                    dist.barrier()
                    torch.cuda.synchronize()
                    comm.allreduce_bucket(bufs[it % 4], bufs[it % 4], scale=1.0, algo=algo)
                    torch.cuda.synchronize()
                    raw = (C.c_uint64 * (256 * 8))()
                    _ffi.check(_ffi.lib().tok_comm_debug_read(comm._h, raw, 256 * 8))
                    a = np.frombuffer(raw, dtype=np.uint64).reshape(256, 8).astype(np.int64)
                    a = a[a[:, 0] > 0]
                    if it >= 3:
                        rows.append(a)
                a = np.stack([r[:min(len(x) for x in rows)] for r in rows])  # it x cta x 8
                t0 = a[:, :, 0].min(axis=1, keepdims=True)
                ph = np.diff(a[:, :, :6], axis=2).astype(np.float64) / 1e3   # us
                stat = dict(stage=np.median(ph[:, :, 0]), barA=np.median(ph[:, :, 1]),
                            reduce=np.median(ph[:, :, 2]), barB=np.median(ph[:, :, 3]),
                            gather=np.median(ph[:, :, 4]),
                            total=float(np.median((a[:, :, 5].max(axis=1) - t0[:, 0]) / 1e3)),
                            start_skew=float(np.median((a[:, :, 0].max(axis=1) - t0[:, 0]) / 1e3)))
                t = torch.tensor([stat[k] for k in sorted(stat)], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                row = dict(zip(sorted(stat), [round(float(x), 2) for x in t.tolist()]))
                row.update(algo=comm.last_algo(), mb=mb, world=world, ctas=ctas, n_ctas=int(a.shape[1]))
                out.append(row)
                if rank == 0:
                    print(json.dumps(row), flush=True)
        comm.close()
    if rank == 0:
        with open(os.path.join(ROOT, "gpurun_out", "phases_n%d.json" % world), "w") as f:
            json.dump(out, f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
