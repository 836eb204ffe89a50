"""Allreduce bus-bandwidth sweep (BASELINE.json config 5): libtok8s algorithms vs NCCL on the same
box.  Launch with torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29511 tools/sweep.py [--max-mb 256] [--dtypes bf16,f32] [--out gpurun_out/sweep]

CUDA-event timing on the launching stream, max over ranks, inputs rotated over > L2-size of buffers.
busbw = S/t * 2(N-1)/N (nccl-tests convention).  Not part of the product.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import ClockSampler  # noqa: E402  (nvidia-smi clocks + throttle reasons during the run)
from torch_on_k8s_b200 import _ffi  # noqa: E402
from torch_on_k8s_b200.comm import Communicator  # noqa: E402

DT = {"bf16": torch.bfloat16, "f32": torch.float32, "f16": torch.float16}


def time_loop(fn, bufs, warmup, iters, stream):
    with torch.cuda.stream(stream):
        for i in range(warmup):
            fn(bufs[i % len(bufs)])
        stream.synchronize()
        dist.barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(iters):
            fn(bufs[i % len(bufs)])
        e1.record(stream)
        stream.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-kb", type=int, default=4)
    ap.add_argument("--max-mb", type=int, default=256)
    ap.add_argument("--dtypes", default="bf16,f32")
    ap.add_argument("--algos", default="2,3,4,0")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/sweep")
    ap.add_argument("--no-nccl", action="store_true")
    ap.add_argument("--zero-copy", action="store_true")
    ap.add_argument("--ctas", default="")  # comma list: rebuild communicator per value
    args = ap.parse_args()

    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream()
    results = []
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    cta_list = [int(x) for x in args.ctas.split(",") if x] or [0]
    for ctas in cta_list:
        if ctas:
            os.environ["TOK_MAX_CTAS"] = str(ctas)
        comm = Communicator("sweep", rank, world, local,
                            rendezvous_path="/tmp/tok8s-sweep-%s-%d" %
                            (os.environ.get("MASTER_PORT", "0"), ctas))
        caps = comm.caps()
        if rank == 0:
            print("caps: multicast=%d staging=%d max_ctas=%d" %
                  (caps.multicast, caps.staging_bytes, caps.max_ctas), flush=True)
        for dname in args.dtypes.split(","):
            dt = DT[dname]
            esz = torch.empty(0, dtype=dt).element_size()
            nbytes = args.min_kb * 1024
            while nbytes <= args.max_mb * (1 << 20):
                count = nbytes // esz
                nbuf = max(2, min(64, (256 << 20) // nbytes))
                bufs = [torch.full((count,), float(rank + 1), dtype=dt, device="cuda")
                        for _ in range(nbuf)]
                iters = args.iters if nbytes <= (64 << 20) else max(10, args.iters // 4)
                row = dict(bytes=nbytes, dtype=dname, world=world, max_ctas=caps.max_ctas)
                for algo in [int(a) for a in args.algos.split(",")]:
                    if algo == 4 and not caps.multicast:
                        continue
                    if algo == 2 and nbytes > caps.staging_bytes // 8:
                        continue

                    def fn(b, algo=algo):
                        comm.allreduce_bucket(b, b, scale=1.0, algo=algo, stream=stream)
                    ms = time_loop(fn, bufs, args.warmup, iters, stream)
                    comm.status()
                    name = _ffi.ALGO_NAMES[algo]
                    row[name + "_us"] = ms * 1e3
                    row[name + "_busbw"] = nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9
                    if algo == 0:
                        row["auto_algo"] = _ffi.ALGO_NAMES[comm.algo_for(nbytes)]
                # zero-copy: the same exchange on buckets that live in the symmetric pool
                # the pool must hold this size's buffer set plus the check buffer on top of what the
                # torch MemPool already caches from the smaller sizes (it does not give segments back)
                pool_ok = (nbuf + 1) * nbytes <= comm.symm_info()[1] - comm.symm_info()[2]
                if args.zero_copy and world > 1 and nbytes >= (1 << 16) and not pool_ok:
                    row["zc_skipped"] = "symmetric pool too small for %d x %d bytes (TOK_SYMM_POOL_MB)" % \
                        (nbuf + 1, nbytes)
                if args.zero_copy and world > 1 and nbytes >= (1 << 16) and pool_ok:
                    zbufs = [comm.symm_empty(count, dt).fill_(float(rank + 1)) for _ in range(nbuf)]
                    zalgos = [(0, "zc_auto"), (3, "zc_two_shot")]
                    if caps.multicast:
                        zalgos.append((4, "zc_nvls"))
                    for algo, name in zalgos:
                        def fnz(b, algo=algo):
                            comm.allreduce_bucket(b, b, scale=1.0, algo=algo, stream=stream)
                        ms = time_loop(fnz, zbufs, args.warmup, iters, stream)
                        comm.status()
                        row[name + "_us"] = ms * 1e3
                        row[name + "_busbw"] = nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9
                        if algo == 0:
                            row["zc_auto_kernel"] = comm.last_algo()
                    # broadcast of the same bytes from rank 0 (parameter / state replication)
                    def fnb(b):
                        comm.broadcast(b, 0, stream=stream)
                    ms = time_loop(fnb, zbufs, args.warmup, iters, stream)
                    comm.status()
                    row["bcast_pool_us"] = ms * 1e3
                    row["bcast_pool_gbs"] = nbytes / (ms * 1e-3) / 1e9
                    row["bcast_kernel"] = comm.last_algo()
                    zb = comm.symm_empty(count, dt).fill_(float(rank + 1))
                    with torch.cuda.stream(stream):
                        comm.allreduce_bucket(zb, zb, scale=1.0, stream=stream)
                        stream.synchronize()
                    row["zc_check"] = bool((zb == float(world * (world + 1) // 2)).all().item())
                    del zbufs, zb
                # correctness spot check: sum of rank+1 (exact in every dtype for world <= 8)
                b = torch.full((count,), float(rank + 1), dtype=dt, device="cuda")
                with torch.cuda.stream(stream):
                    comm.allreduce_bucket(b, b, scale=1.0, stream=stream)
                    stream.synchronize()
                row["check"] = bool((b == float(world * (world + 1) // 2)).all().item())
                if not args.no_nccl:
                    def fn2(b):
                        dist.all_reduce(b)
                    ms = time_loop(fn2, bufs, args.warmup, iters, stream)
                    row["nccl_us"] = ms * 1e3
                    row["nccl_busbw"] = nbytes / (ms * 1e-3) * 2 * (world - 1) / world / 1e9

                    def fn3(b):
                        dist.broadcast(b, 0)
                    ms = time_loop(fn3, bufs, args.warmup, iters, stream)
                    row["nccl_bcast_us"] = ms * 1e3
                results.append(row)
                if rank == 0:
                    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v)
                                      for k, v in row.items()}), flush=True)
                    # written after every row: a run that dies at the largest sizes keeps its rows
                    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
                    with open("%s_n%d.json" % (args.out, world), "w") as f:
                        json.dump(dict(world=world, when=time.time(), partial=True,
                                       gpu=torch.cuda.get_device_name(0), rows=results), f, indent=1)
                del bufs, b
                nbytes *= 4 if nbytes < (1 << 20) else 2
        comm.close()
        dist.barrier()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open("%s_n%d.json" % (args.out, world), "w") as f:
            json.dump(dict(world=world, when=time.time(), clocks=sampler.stop(),
                           gpu=torch.cuda.get_device_name(0), rows=results), f, indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
