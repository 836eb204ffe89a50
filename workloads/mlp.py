"""BASELINE.json config 0: 2-layer MLP on synthetic MNIST-shaped data (Linear(784,128)-ReLU-
Linear(128,10)); 101,770 parameters, one fp32 DDP bucket of 407,080 bytes."""
import torch
import torch.nn as nn


def mlp(seed: int = 0) -> nn.Sequential:
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = nn.Sequential(nn.Linear(784, 128), nn.ReLU(), nn.Linear(128, 10))
    torch.random.set_rng_state(g)
    return m


def batch(rank: int, n: int = 64):
    """x ~ N(0,1)[n,784], y ~ U{0..9}, from Generator(seed = 7 + rank)."""
    gen = torch.Generator().manual_seed(7 + rank)
    x = torch.randn(n, 784, generator=gen)
    y = torch.randint(0, 10, (n,), generator=gen)
    return x, y
