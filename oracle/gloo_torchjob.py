"""ORACLE / CPU BASELINE — TEST INFRASTRUCTURE ONLY (see oracle/allreduce_oracle.py header).

The reference-style torchjob: N local CPU processes wired exactly as the reference operator would
wire its pods (TorchJobReconciler.SetClusterSpec, controllers/train/torchjob_controller.go:314-449:
master RANK 0, worker i RANK i+1, WORLD_SIZE = masters + workers, MASTER_PORT 23456 by default,
PYTHONUNBUFFERED=0), `init_process_group("gloo", "env://")`, DistributedDataParallel with PyTorch's
own reducer and gloo allreduce.  This IS the reference's hot-path implementation (the third-party
dependency the operator delegates to), run live; it is used
  * by tests/golden/make_golden.py to generate the golden vectors that pin oracle/allreduce_oracle.py,
  * by the CPU world_size-2 tests,
  * by bench.py's `cpu_baseline` leg and `--impl reference` arm (timed, never shipped).
"""
from __future__ import annotations

import json
import multiprocessing as mp
import os
import socket
import sys
import time
import traceback
from typing import Dict, List, Optional

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def effective_cores() -> int:
    """Host cores this container may actually use: min(affinity, cgroup cpu.max quota).  The GPU box
    reports 128 logical CPUs but a 16-CPU quota; oversubscribing it makes OpenMP spin for minutes."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except Exception:  # noqa: BLE001
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, (q + p - 1) // p))
        except Exception:  # noqa: BLE001
            pass
    return max(1, n)


def replica_env(job: str, task_type: str, index: int, num_workers: int, port: int = 23456,
                local_master_addr: bool = True) -> Dict[str, str]:
    """Oracle restatement of the env half of SetClusterSpec (torchjob_controller.go:338-350,
    394-446) for a 1-master + num_workers job."""
    tt = task_type.lower()
    master_addr = ("%s-master-0" % job).replace("/", "-")
    if tt == "master":
        if index != 0:
            raise ValueError("invalid config: There should be only a single master with index=0")
        rank = 0
        if local_master_addr:  # feature gate TorchLocalMasterAddr, default on
            master_addr = "localhost"
    else:
        rank = index + 1
    return {"MASTER_PORT": str(port), "MASTER_ADDR": master_addr, "RANK": str(rank),
            "PYTHONUNBUFFERED": "0", "WORLD_SIZE": str(1 + num_workers)}


def _build(workload: str):
    import torch
    if workload == "mlp":
        from workloads.mlp import mlp
        return mlp(0)
    if workload == "resnet50":
        # stock torchvision model: the reference arm shares no model code with the product side
        from torchvision.models import resnet50
        torch.manual_seed(0)
        return resnet50(num_classes=1000)
    raise ValueError(workload)


def _batch(workload: str, rank: int, n: int, dtype):
    import torch
    if workload == "mlp":
        from workloads.mlp import batch
        return batch(rank, n)
    gen = torch.Generator().manual_seed(1234 + rank)
    x = torch.randn(n, 3, 224, 224, generator=gen).to(dtype).contiguous(
        memory_format=torch.channels_last)
    y = torch.randint(0, 1000, (n,), generator=gen)
    return x, y


# Everything a launcher (torchrun, a test runner, a scheduler) may have left in the environment
# that changes how init_process_group("gloo", "env://") or OpenMP behave.  The reference writes ONLY
# the SetClusterSpec variables into a container (torchjob_controller.go:394-446): nothing else is
# inherited.  With TORCHELASTIC_USE_AGENT_STORE set, env:// waits for an agent store nobody started.
_SCRUB_PREFIXES = ("TORCHELASTIC_", "GROUP_", "ROLE_", "LOCAL_", "NCCL_", "TORCH_NCCL_", "OMP_",
                   "MKL_", "KMP_", "GOMP_", "TORCHINDUCTOR_", "PET_", "TOK8S_", "TOK_")
_SCRUB_KEYS = ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "PYTHONUNBUFFERED")


def clean_env(extra: Dict[str, str]) -> Dict[str, str]:
    """The environment of one replica container: the caller's environment minus everything a
    launcher injected, plus the SetClusterSpec variables in `extra`."""
    env = {k: v for k, v in os.environ.items()
           if not k.startswith(_SCRUB_PREFIXES) and k not in _SCRUB_KEYS}
    env["PYTHONPATH"] = ROOT + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
    env.update(extra)
    return env


def _replica(rank, world, env, workload, steps, warmup, batch, threads, dtype_name, dump, q):
    try:
        for k in list(os.environ):
            if k.startswith(_SCRUB_PREFIXES) or k in _SCRUB_KEYS:
                del os.environ[k]
        os.environ.update(env)
        # single box: the master's service name resolves to loopback
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        import torch
        import torch.distributed as dist
        from torch.nn.parallel import DistributedDataParallel as DDP
        torch.set_num_threads(max(1, threads))
        dist.init_process_group("gloo", init_method="env://")
        assert dist.get_rank() == rank and dist.get_world_size() == world
        dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[dtype_name]
        model = _build(workload).to(dtype)
        if workload == "resnet50":
            model = model.to(memory_format=torch.channels_last)
        ddp = DDP(model, bucket_cap_mb=25)
        record: List[dict] = []
        if dump:
            def hook(state, bucket):
                pre = bucket.buffer().clone()
                fut = dist.all_reduce(bucket.buffer().div_(world), async_op=True).get_future()

                def done(f):
                    out = f.value()[0]
                    record.append(dict(index=bucket.index(), pre=pre, post=out.clone()))
                    return out
                return fut.then(done)
            ddp.register_comm_hook(None, hook)
        # same optimizer as the product arm of bench.py for ResNet-50; plain SGD for the MLP goldens
        opt = torch.optim.SGD(ddp.parameters(), lr=0.01,
                              momentum=0.9 if workload == "resnet50" else 0.0)
        lossf = torch.nn.CrossEntropyLoss()
        x, y = _batch(workload, rank, batch, dtype)
        losses = []
        t0 = None
        for it in range(warmup + steps):
            if it == warmup:
                dist.barrier()
                t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            loss = lossf(ddp(x).float(), y)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        dist.barrier()
        dt = time.perf_counter() - t0
        out = dict(rank=rank, seconds=dt, steps=steps, losses=losses, threads=torch.get_num_threads())
        if dump:
            import numpy as np
            os.makedirs(dump, exist_ok=True)
            arrays = {}
            for k, r in enumerate(record):
                arrays["it%03d_b%d_pre" % (k, r["index"])] = r["pre"].float().numpy()
                if rank == 0:  # every rank holds the same bits after the allreduce
                    arrays["it%03d_b%d_post" % (k, r["index"])] = r["post"].float().numpy()
            with torch.no_grad():
                arrays["argmax"] = ddp(x).float().argmax(1).numpy()
                arrays["final_flat"] = torch.cat([p.detach().float().flatten()
                                                  for p in ddp.parameters()]).numpy()
            np.savez_compressed(os.path.join(dump, "rank%d.npz" % rank), **arrays)
        dist.destroy_process_group()
        q.put((rank, "ok", out))
    except Exception:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))


def run(workload: str = "mlp", world: int = 2, steps: int = 5, warmup: int = 1, batch: int = 64,
        threads: Optional[int] = None, dtype: str = "f32", dump: Optional[str] = None,
        job: str = "torchjob", timeout: float = 1800.0) -> dict:
    """Run the gloo/CPU torchjob with `world` replicas (1 master + world-1 workers) and return
    {"seconds", "steps", "images_per_sec", "cores", "threads_per_replica", "losses"}."""
    cores = effective_cores()
    # cores are split between the replicas; more than 16 threads per replica only adds OpenMP
    # spinning at these batch sizes (measured: 96 threads 23 img/s vs 16 threads 74 img/s)
    threads = threads or max(1, min(cores // world, 16))
    port = free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = []
    for r in range(world):
        env = replica_env(job, "master" if r == 0 else "worker", 0 if r == 0 else r - 1, world - 1,
                          port)
        p = ctx.Process(target=_replica, args=(r, world, env, workload, steps, warmup, batch,
                                               threads, dtype, dump, q), daemon=True)
        p.start()
        procs.append(p)
    results, errors = {}, []
    deadline = time.time() + timeout
    while len(results) + len(errors) < world and time.time() < deadline:
        try:
            rank, status, payload = q.get(timeout=5)
        except Exception:  # noqa: BLE001
            if any(p.exitcode not in (None, 0) for p in procs):
                errors.append((-1, "a replica process died: %s" % [p.exitcode for p in procs]))
                break
            continue
        (results.__setitem__(rank, payload) if status == "ok" else errors.append((rank, payload)))
    for p in procs:
        p.join(timeout=5)
        if p.is_alive():
            p.kill()
    if errors or len(results) < world:
        raise RuntimeError("gloo torchjob failed: %s" % (errors[:1] or "timeout"))
    sec = max(r["seconds"] for r in results.values())
    return dict(workload=workload, world=world, steps=steps, warmup=warmup, batch_per_replica=batch,
                seconds=sec, images_per_sec=world * batch * steps / sec, cores=cores,
                threads_per_replica=results[0]["threads"], losses=results[0]["losses"],
                dtype=dtype)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="mlp")
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dtype", default="f32")
    a = ap.parse_args()
    print(json.dumps(run(a.workload, a.world, a.steps, a.warmup, a.batch, dtype=a.dtype)))
