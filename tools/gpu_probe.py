"""First-contact probe for a gpurun box: capabilities, single-GPU multi-replica parity in both
process and thread mode, and a first HBM number for the world==1 fused scale/cast kernel.
Writes gpurun_out/probe.json.  Not part of the product."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

import torch  # noqa: E402

import harness  # noqa: E402


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout
    except Exception as e:  # noqa: BLE001
        return "ERR %s" % e


def main():
    out = {"t0": time.time()}
    out["nvidia_smi"] = sh("nvidia-smi --query-gpu=index,name,memory.total,clocks.sm,clocks.max.sm "
                           "--format=csv")
    out["topo"] = sh("nvidia-smi topo -m")
    out["nproc"] = os.cpu_count()
    out["cpu"] = sh("lscpu | head -20")
    n = torch.cuda.device_count()
    out["device_count"] = n
    print("devices:", n, flush=True)
    print(out["nvidia_smi"], flush=True)
    print(out["topo"], flush=True)

    # ---- world == 1: fused scale/cast, correctness + bandwidth ---------------------------------
    from torch_on_k8s_b200.comm import Communicator
    comm = Communicator("probe1", 0, 1, 0, rendezvous_path="/tmp/tok8s-probe1")
    caps = comm.caps()
    out["caps_w1"] = {f[0]: getattr(caps, f[0]) for f in caps._fields_}
    print("caps:", out["caps_w1"], flush=True)
    res = []
    for (din, dw, dout) in [("bf16", "bf16", "bf16"), ("f32", "bf16", "f32"), ("f32", "f32", "f32")]:
        for count in ((4098000 // 2,) if os.environ.get("PROBE_SKIP_SHARED") else (4098000 // 2, 28878848 // 2, 1 << 28)):
            x = torch.randn(count, device="cuda").to(harness.torch_dtype(din))
            y = torch.empty(count, device="cuda", dtype=harness.torch_dtype(dout))
            for _ in range(3):
                comm.allreduce_bucket(x, y, scale=0.5, wire_dtype=harness.torch_dtype(dw))
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            iters = 20
            ev[0].record()
            for _ in range(iters):
                comm.allreduce_bucket(x, y, scale=0.5, wire_dtype=harness.torch_dtype(dw))
            ev[1].record()
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1]) / iters
            nbytes = count * (x.element_size() + y.element_size())
            want = (x.float() * 0.5).to(harness.torch_dtype(dw)).to(harness.torch_dtype(dout))
            ok = bool(torch.equal(want, y))
            r = dict(triple=[din, dw, dout], count=count, ms=ms, gbs=nbytes / ms / 1e6, exact=ok)
            print("local", r, flush=True)
            res.append(r)
            del x, y, want
    out["local"] = res
    comm.close()

    # ---- multi-replica parity on a shared GPU ----------------------------------------------------
    env = {"TOK_MAX_CTAS": "16", "TOK_BARRIER_TIMEOUT_MS": "30000", "TOK_STAGING_MB": "32"}
    out["shared"] = []
    for mode in (() if os.environ.get("PROBE_SKIP_SHARED") else ("thread", "proc")):
        for world in (2, 4, 8):
            cases = harness.standard_cases(world, algos=(2, 3), quick=True)
            t0 = time.time()
            try:
                rs = harness.launch(world, cases, devices=[0] * world, mode=mode, timeout=420,
                                    env=env, job="probe-%s-%d" % (mode, world))
                s = harness.summarize(rs)
                s.update(mode=mode, world=world, sec=time.time() - t0, ncases=len(cases))
                ms = [r["ms"] for r in rs[0] if "ms" in r]
                s["median_ms"] = sorted(ms)[len(ms) // 2] if ms else None
            except Exception as e:  # noqa: BLE001
                s = dict(mode=mode, world=world, error=str(e)[-2000:], sec=time.time() - t0)
            print("shared", json.dumps(s)[:3000], flush=True)
            out["shared"].append(s)
            with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
                json.dump(out, f, indent=1, default=str)

    # ---- one replica per GPU when the box has several ---------------------------------------------
    if n >= 2:
        out["multi"] = []
        for world in sorted({2, min(n, 4), n}):
            cases = harness.standard_cases(world, algos=(2, 3, 4), quick=False)
            t0 = time.time()
            try:
                rs = harness.launch(world, cases, devices=list(range(world)), mode="proc",
                                    timeout=420, job="probe-multi-%d" % world)
                s = harness.summarize(rs)
                s.update(world=world, sec=time.time() - t0, ncases=len(cases), tail=rs[0][-1])
            except Exception as e:  # noqa: BLE001
                s = dict(world=world, error=str(e)[-2000:], sec=time.time() - t0)
            print("multi", json.dumps(s)[:3000], flush=True)
            out["multi"].append(s)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1, default=str)
    print("probe done", flush=True)


if __name__ == "__main__":
    main()
