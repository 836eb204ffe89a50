#!/usr/bin/env python
"""bench.py — BASELINE.json's metric: ResNet-50 synthetic-image throughput (images/s) of a
data-parallel TorchJob at N worker-replica GPUs, with the allreduce's achieved bus bandwidth against
the NVLink roofline, next to the reference-style gloo/CPU torchjob on the box's host cores.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one synthetic batch: forward + backward of ResNet-50
(bf16 parameters/gradients, channels_last, per-GPU batch 256, DDP bucket cap 25 MB -> 3 buckets of
4.1/28.9/18.1 MB) whose gradient buckets are averaged across replicas by libtok8s' fused
cast/scale/allreduce kernels (through torch_on_k8s_b200's DDP comm hook), + SGD step.
Timing: CUDA events, barrier + synchronize on both sides, MAX over ranks; rank 0 prints ONE JSON
line.  `value` has inputs resident in HBM; `e2e` copies every step's batch from pinned host memory
and reads the loss back.  The step's working set (GBs of activations) is far larger than the
126 MB L2, so no explicit L2 flush is needed between iterations (stated in config.l2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "resnet50_ddp_images_per_sec"
UNIT = "images/s"
NVLINK_PEAK_GBS = 900.0  # nominal per direction per GPU (BASELINE.md §2 fixes this denominator)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (profiling recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference-style gloo/CPU torchjob (oracle/gloo_torchjob.py), timed on the host
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0  # rank 0 alone runs the CPU job; other torchrun ranks exit without work
    from oracle import gloo_torchjob
    n = args.gpus
    cores = gloo_torchjob.effective_cores()
    batch = args.ref_batch
    t0 = time.time()
    res = gloo_torchjob.run("resnet50", world=n, steps=args.steps, warmup=args.warmup, batch=batch,
                            dtype="bf16", job="bench-ref")
    value = res["images_per_sec"]
    sample = ("%d timed + %d warm-up steps of the same workload at per-replica batch %d (bf16, "
              "gloo/TCP loopback, %d replicas x %d threads)" %
              (args.steps, args.warmup, batch, n, res["threads_per_replica"]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": res["seconds"] / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "ResNet-50 bf16 synthetic 224x224 images, reference-style torchjob "
                               "(1 master + %d workers, env per SetClusterSpec, gloo CPU backend, "
                               "DDP bucket 25MB), per-replica batch %d" % (n - 1, batch),
                   "global_batch": batch * n, "parallelism": "dp%d" % n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback "
                         "(use --impl reference for the gloo/CPU torchjob)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch ourselves under torchrun
            from oracle.gloo_torchjob import free_port
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
                   "--master-port", str(free_port()), os.path.abspath(__file__)] + sys.argv[1:]
            return subprocess.call(cmd)
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if "MASTER_PORT" not in os.environ:
        from oracle.gloo_torchjob import free_port
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(free_port())
    os.environ.setdefault("TOK8S_JOB", "bench-resnet50")

    from torch_on_k8s_b200.worker import init_replica
    from workloads.resnet50 import resnet50

    rep = init_replica(device=local)
    dev = rep.device
    torch.backends.cudnn.benchmark = True
    B = args.batch
    torch.manual_seed(0)
    model = resnet50().to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    ddp, hook = rep.wrap(model, bucket_cap_mb=25, record_events=True)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.01, momentum=0.9)
    lossf = torch.nn.CrossEntropyLoss()

    gen = torch.Generator().manual_seed(1234 + rank)
    n_host = 4  # rotating pinned host batches for the e2e leg
    host_x = [torch.randn(B, 3, 224, 224, generator=gen).to(torch.bfloat16)
              .contiguous(memory_format=torch.channels_last).pin_memory() for _ in range(n_host)]
    host_y = [torch.randint(0, 1000, (B,), generator=gen).pin_memory() for _ in range(n_host)]
    dev_x = host_x[0].to(dev, non_blocking=True)
    dev_y = host_y[0].to(dev, non_blocking=True)

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        loss = lossf(ddp(x).float(), y)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return float(t.item())

    # ---- warm-up (also lets DDP rebuild its buckets after iteration 1) ---------------------------
    W = max(args.warmup, 3)
    for _ in range(W):
        step(dev_x, dev_y)
    barrier()
    hook.drain_events()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = rep.comm.launches()
    ms_dev = timed(lambda i: step(dev_x, dev_y), args.steps)
    launches = rep.comm.launches() - launches0
    ev = hook.drain_events()

    # ---- end-to-end leg: host batch -> device every step, loss read back every step --------------
    h2d = host_x[0].numel() * host_x[0].element_size() + host_y[0].numel() * host_y[0].element_size()

    def e2e_step(i):
        x = host_x[i % n_host].to(dev, non_blocking=True)
        y = host_y[i % n_host].to(dev, non_blocking=True)
        loss = step(x, y)
        return float(loss.detach().float().item())  # device->host read of the step's result

    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    hook.drain_events()
    rep.comm.status()

    # ---- roofline of the dominant kernel of OUR path (the bucket allreduce) ----------------------
    peaks, peak_kind = measured_peaks()
    wire_bytes = sum(e[0] for e in ev)
    bucket_bytes = sum(e[1] for e in ev)
    ar_ms = sum(e[2] for e in ev)
    n_launch = max(len(ev), 1)
    stats = torch.tensor([wire_bytes, ar_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    ar_ms_max = float(stats[1].item())
    per_step_bytes = wire_bytes / args.steps
    if world == 1:
        # fused scale/cast only: S_in + S_out against HBM
        achieved = (2.0 * bucket_bytes) / (ar_ms_max * 1e-3) / 1e9 if ar_ms_max > 0 else 0.0
        roof = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None,
                "kernel": "tok::local_kernel<bf16,bf16,bf16>",
                "algorithmic_bytes_per_launch": 2.0 * bucket_bytes / n_launch,
                "avg_launch_us": ar_ms_max * 1e3 / n_launch, "peak_source": peak_kind + " hbm_gbs",
                "traffic_note": "ncu --set full of the same kernel at 2x512 MiB: dram read+write "
                                "1015 MB per launch vs 1074 MB algorithmic, 6.04 TB/s "
                                "(profiles/r01_ncu_local_kernel_raw.csv); not captured at the "
                                "bucket sizes of this run, hence traffic=null",
                "note": "world=1 degenerates to the fused scale/cast copy: bytes = S_in + S_out"}
    else:
        algbw = wire_bytes / (ar_ms_max * 1e-3) / 1e9 if ar_ms_max > 0 else 0.0
        busbw = algbw * 2.0 * (world - 1) / world
        roof = {"bound": "nvlink", "achieved": busbw, "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
                "frac": busbw / NVLINK_PEAK_GBS, "traffic": None,
                "kernel": "tok::{one_shot,two_shot,nvls}_kernel<bf16,bf16,bf16>",
                "algbw_gbs": algbw, "algorithmic_bytes_per_launch": wire_bytes / n_launch,
                "avg_launch_us": ar_ms_max * 1e3 / n_launch,
                "peak_source": "nominal NVLink5 900 GB/s per direction (BASELINE.md §2); measured "
                               "peer copy on this pool is 770 GB/s",
                "note": "busbw = S/t * 2(N-1)/N over the bucket launches inside the timed steps "
                        "(includes waiting for the slowest replica to reach the bucket); NVLS may "
                        "exceed 1.0"}

    # ---- the same buckets, exchanged back to back with nothing else on the GPU (kernel quality
    #      without the wait for the slowest replica's backward); buffers rotate over > L2 bytes ------
    sizes = [e[0] for e in ev[:max(1, len(ev) // args.steps)]]
    iso = None
    if sizes:
        el = torch.empty(0, dtype=torch.bfloat16).element_size()
        nsets = max(2, int((160 << 20) / max(sum(sizes), 1)) + 1)
        zero_copy = hook.zero_copy_buckets > 0   # same placement as DDP's own buckets

        def mk(sz):
            t = rep.comm.symm_empty(sz // el, torch.bfloat16) if zero_copy else \
                torch.empty(sz // el, device=dev, dtype=torch.bfloat16)
            return t.normal_()
        sets = [[mk(sz) for sz in sizes] for _ in range(min(nsets, 8))]
        cstream = torch.cuda.Stream(device=dev)

        def iso_step(i):
            for t in sets[i % len(sets)]:
                rep.comm.allreduce_bucket(t, t, scale=1.0 / world, stream=cstream)

        with torch.cuda.stream(cstream):
            for i in range(5):
                iso_step(i)
            cstream.synchronize()
            barrier()
            i0 = torch.cuda.Event(enable_timing=True)
            i1 = torch.cuda.Event(enable_timing=True)
            iters = 30
            i0.record(cstream)
            for i in range(iters):
                iso_step(i)
            i1.record(cstream)
            cstream.synchronize()
        tt = torch.tensor([i0.elapsed_time(i1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        iso_ms = float(tt.item())
        iso_bytes = float(sum(sizes) * iters)
        if world == 1:
            ach = 2.0 * iso_bytes / (iso_ms * 1e-3) / 1e9
            iso = {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "frac": ach / peaks["hbm_gbs"], "avg_launch_us": iso_ms * 1e3 / (iters * len(sizes))}
        else:
            ach = iso_bytes / (iso_ms * 1e-3) / 1e9 * 2.0 * (world - 1) / world
            iso = {"bound": "nvlink", "achieved": ach, "peak": NVLINK_PEAK_GBS, "unit": "GB/s",
                   "frac": ach / NVLINK_PEAK_GBS, "avg_launch_us": iso_ms * 1e3 / (iters * len(sizes))}
        iso["bucket_bytes"] = sizes
        iso["zero_copy"] = bool(zero_copy)
        iso["note"] = ("the step's buckets exchanged back to back on an otherwise idle GPU, inputs "
                       "rotated over >126 MB so they are not L2 resident")
        rep.comm.status()

    if rank != 0:
        rep.close()
        return 0

    images = args.steps * B * world
    value = images / (ms_dev * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": W, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "ResNet-50 bf16 synthetic 224x224 images (random-init weights), "
                               "per-GPU batch %d, DDP bucket 25MB (4.1/28.9/18.1 MB at iteration 0, "
                               "rebuilt by DDP into 28.3/22.9 MB; 51.1 MB/step/replica), SGD "
                               "momentum, channels_last" % B,
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "step working set (GBs of activations) is larger than the 126 MB L2; "
                         "no explicit flush",
                   "allreduce_bytes_per_step": per_step_bytes},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches * world),
        "zero_copy_buckets": bool(hook.zero_copy_buckets > 0),
        "roofline": roof,
        "roofline_isolated": iso,
    }
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on the host cores: the reference-style gloo job
        from oracle import gloo_torchjob
        cb = gloo_torchjob.run("resnet50", world=1, steps=args.cpu_steps, warmup=1,
                               batch=args.ref_batch, dtype="bf16", job="bench-cpu")
        line["cpu_baseline"] = {
            "value": cb["images_per_sec"], "unit": UNIT, "cores": cb["cores"], "kind": "reference",
            "sample": "%d timed steps (+1 warm-up) of the same ResNet-50 bf16 step at batch %d on "
                      "the host CPU: torch DDP + gloo configured per SetClusterSpec, %d threads" %
                      (args.cpu_steps, args.ref_batch, cb["threads_per_replica"])}
    print(json.dumps(line), flush=True)
    rep.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch (BASELINE config: 256)")
    ap.add_argument("--ref-batch", type=int, default=8,
                    help="per-replica batch of the CPU arm (bounded sample of the same workload)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
