"""Replica body for the CPU controller test: a reference-style gloo MLP worker that consumes ONLY the
env contract the controller injects (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT) and writes its
losses to $OUT_DIR/<replica>.json.  Test infrastructure."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

from workloads.mlp import batch, mlp  # noqa: E402

if os.environ.get("FAIL_ONCE") and not os.path.exists(os.environ["FAIL_ONCE"]) and \
        os.environ.get("TOK8S_TASK_TYPE") == os.environ.get("FAIL_TYPE", "Master"):
    open(os.environ["FAIL_ONCE"], "w").write("x")
    os._exit(int(os.environ.get("FAIL_CODE", "137")))

torch.set_num_threads(1)
dist.init_process_group("gloo", init_method="env://")
rank, world = dist.get_rank(), dist.get_world_size()
assert rank == int(os.environ["RANK"]) and world == int(os.environ["WORLD_SIZE"])
ddp = DDP(mlp(0))
opt = torch.optim.SGD(ddp.parameters(), lr=0.01)
x, y = batch(rank, 64)
losses = []
for _ in range(int(os.environ.get("STEPS", "2"))):
    opt.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(ddp(x), y)
    loss.backward()
    opt.step()
    losses.append(float(loss.detach()))
with open(os.path.join(os.environ["OUT_DIR"], os.environ["TOK8S_REPLICA"] + ".json"), "w") as f:
    json.dump(dict(rank=rank, world=world, losses=losses, env={k: os.environ.get(k) for k in
                   ("RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "TOK8S_GPU", "TOK8S_RDZV",
                    "PYTHONUNBUFFERED")}), f)
dist.destroy_process_group()
