"""Replica entry point for TorchJob manifests run by the single-box controller
(`python -m torch_on_k8s_b200.controller samples/…yaml`): the training script that would live in the
user's container.  It reads the reference's env contract (RANK / WORLD_SIZE / MASTER_*) through
init_replica(), wraps the model so that every gradient bucket goes through libtok8s, and prints the
progress line the torchelastic controller parses (controllers/train/torchelastic/observation.go:54-76):
"Epoch: [e][ b/N]\\tTime  0.ddd ( 0.ddd)\\tData ...\\tLoss ...\\tAcc@1 ...\\tAcc@5  dd.dd ( dd.dd)".
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build(name: str):
    import torch
    if name == "mlp":
        from workloads.mlp import mlp
        return mlp(0), (784,), 10, torch.float32
    if name == "resnet50":
        from workloads.resnet50 import resnet50
        torch.manual_seed(0)
        return resnet50(), (3, 224, 224), 1000, torch.bfloat16
    raise SystemExit("unknown --model %s" % name)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--log-every", type=int, default=10)
    a = ap.parse_args()

    import torch
    from torch_on_k8s_b200.sampler import ReplicaSampler
    from torch_on_k8s_b200.worker import init_replica, report_metric

    rep = init_replica()
    model, shape, classes, dtype = build(a.model)
    model = model.to(rep.device).to(dtype)
    if len(shape) == 3:
        model = model.to(memory_format=torch.channels_last)
    ddp, hook = rep.wrap(model, record_events=(rep.rank == 0))
    opt = torch.optim.SGD(ddp.parameters(), lr=0.01, momentum=0.9)
    sampler = ReplicaSampler(a.steps * a.batch * rep.world, rep.world, rep.rank, seed=0)
    gen = torch.Generator(device=rep.device).manual_seed(1234 + rep.rank)
    t_prev = time.time()
    for step, _ in zip(range(a.steps), iter(sampler)):
        x = torch.randn((a.batch,) + shape, device=rep.device, generator=gen).to(dtype)
        if len(shape) == 3:
            x = x.contiguous(memory_format=torch.channels_last)
        y = torch.randint(0, classes, (a.batch,), device=rep.device, generator=gen)
        opt.zero_grad(set_to_none=True)
        out = ddp(x).float()
        loss = torch.nn.functional.cross_entropy(out, y)
        loss.backward()
        opt.step()
        if (step + 1) % a.log_every == 0 and rep.rank <= 1:
            torch.cuda.synchronize()
            now = time.time()
            lat = (now - t_prev) / a.log_every
            t_prev = now
            acc = float((out.argmax(1) == y).float().mean()) * 100
            print("Epoch: [0][%4d/%d]\tTime %6.3f (%6.3f)\tData  0.000 ( 0.000)\tLoss %.4e\t"
                  "Acc@1 %6.2f (%6.2f)\tAcc@5 %6.2f (%6.2f)" %
                  (step + 1, a.steps, lat, lat, float(loss), acc, acc, acc, acc), flush=True)
            if rep.rank == 0 and rep.world > 1:
                ev = hook.drain_events()       # [(wire bytes, bucket bytes, exchange ms, wait ms)]
                ms = sum(e[2] for e in ev)
                if ms > 0:                      # nccl-tests convention: S/t * 2(N-1)/N
                    report_metric(busbw_gbps=sum(e[0] for e in ev) / (ms * 1e-3) / 1e9 *
                                  2.0 * (rep.world - 1) / rep.world)
    torch.cuda.synchronize()
    rep.comm.status()
    rep.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
