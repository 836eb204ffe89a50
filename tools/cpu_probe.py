"""Probe the GPU box's host CPU budget (cgroup quota, affinity) and time the CPU ResNet-50 step at
several thread counts / dtypes, to size bench.py's cpu_baseline leg. Not part of the product."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for p in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(p, open(p).read().strip())
    except Exception as e: print(p, "n/a")
print(torch.__config__.parallel_info())
from workloads.resnet50 import resnet50
for dtype in (torch.bfloat16, torch.float32):
    for nt in (8, 16, 32, 64, 128):
        torch.set_num_threads(nt)
        m = resnet50().to(dtype).to(memory_format=torch.channels_last)
        x = torch.randn(8, 3, 224, 224).to(dtype).contiguous(memory_format=torch.channels_last)
        y = torch.randint(0, 1000, (8,))
        lossf = torch.nn.CrossEntropyLoss()
        ts = []
        for it in range(3):
            t0 = time.perf_counter()
            m.zero_grad(set_to_none=True)
            lossf(m(x).float(), y).backward()
            ts.append(time.perf_counter() - t0)
            if ts[-1] > 30: break
        print(json.dumps(dict(dtype=str(dtype), threads=nt, step_s=[round(t, 3) for t in ts])), flush=True)
        if min(ts) > 20: break
