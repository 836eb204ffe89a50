"""Drives the world==1 fused scale/cast kernel on ResNet-50-bucket-sized and large buffers so that
`ncu --set full -k regex:local_kernel` can capture it in isolation. Not part of the product."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch_on_k8s_b200.comm import Communicator
c = Communicator("prof", 0, 1, 0, rendezvous_path="/tmp/tok8s-prof")
for n, dt, wire in ((28878848 // 2, torch.bfloat16, None), (1 << 28, torch.bfloat16, None),
                    (1 << 26, torch.float32, torch.bfloat16)):
    x = torch.randn(n, device="cuda").to(dt)
    for _ in range(3):
        c.allreduce_bucket(x, x, scale=0.5, wire_dtype=wire)
    torch.cuda.synchronize()
c.close()
