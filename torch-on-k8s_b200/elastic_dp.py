"""ElasticDataParallel — data-parallel gradient averaging WITHOUT torch.distributed.

DistributedDataParallel cannot change its world size: the reference therefore restarts every stale
pod when a job is rescaled (controllers/train/elastic_scale.go:210-397, WORLD_SIZE re-read from an
annotation).  Here the peer group re-forms in place (tok_comm_reform / tok_comm_join), so the
training process keeps running; what it needs is a gradient-bucket layer that only depends on the
communicator:

  * gradients live in flat buckets allocated from the replica's SYMMETRIC POOL
    (tok_comm_symm_alloc — same allocation sequence on every replica => same offsets), parameters'
    `.grad` are views into them (DDP's gradient_as_bucket_view layout: reverse parameter order,
    first bucket 1 MiB then `bucket_cap_mb`, torch/nn/parallel/distributed.py:831-853);
  * reduce_grads() is one zero-copy tok_allreduce_bucket per bucket (scale 1/world fused);
  * reform()/sync_params() handle elastic add/drop: survivors keep everything, a joiner builds the
    same bucket sequence and receives the parameters through the same kernels.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from ._ffi import check, lib
from .comm import Communicator


class _RawCuda:
    """Exposes a raw device pointer through __cuda_array_interface__ so that torch can view it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1",
                                         "data": (ptr, False), "version": 2}


def symm_tensor(comm: Communicator, numel: int, dtype: torch.dtype) -> torch.Tensor:
    """A flat tensor inside the replica's symmetric pool (bump-allocated, collective-consistent)."""
    nbytes = numel * torch.empty(0, dtype=dtype).element_size()
    ptr = C.c_void_p()
    check(lib().tok_comm_symm_alloc(comm._h, max(nbytes, 16), C.byref(ptr)))
    raw = torch.as_tensor(_RawCuda(ptr.value, max(nbytes, 16)), device=torch.device("cuda", comm.device))
    return raw[:nbytes].view(dtype)


def bucket_assignment(sizes_bytes, keys, caps):
    """Gradient-bucket assignment of torch's Reducer (compute_bucket_assignment_by_size, reached from
    torch/nn/parallel/distributed.py:1183-1275 — SURVEY.md §8 row a10): walk the tensors in the given
    order, one open bucket per (dtype, device) key; a bucket is closed as soon as its size REACHES
    its limit; the limits advance through `caps` per key ([first 1 MiB, then bucket_cap]); buckets
    are returned ordered by their smallest tensor index.  Pure function (unit-tested on CPU against
    torch.distributed._compute_bucket_assignment_by_size)."""
    open_b, cap_pos, out = {}, {}, []
    for i, (nb, k) in enumerate(zip(sizes_bytes, keys)):
        idx, size = open_b.get(k, ([], 0))
        idx = idx + [i]
        size += nb
        pos = cap_pos.get(k, 0)
        if size >= caps[pos]:
            out.append(idx)
            open_b[k] = ([], 0)
            cap_pos[k] = min(pos + 1, len(caps) - 1)
        else:
            open_b[k] = (idx, size)
    out += [idx for idx, _ in open_b.values() if idx]
    out.sort(key=min)
    return out


class ElasticDataParallel(torch.nn.Module):
    def __init__(self, module: torch.nn.Module, comm: Communicator, *, bucket_cap_mb: int = 25,
                 first_bucket_mb: int = 1, algo: int = 0):
        super().__init__()
        self.module = module
        self.comm = comm
        self.algo = algo
        self.buckets: List[torch.Tensor] = []
        self._assign(bucket_cap_mb << 20, first_bucket_mb << 20)

    def _assign(self, cap: int, first_cap: int) -> None:
        params = [p for p in self.module.parameters() if p.requires_grad][::-1]  # reverse order
        groups = bucket_assignment([p.numel() * p.element_size() for p in params],
                                   [(p.dtype, p.device) for p in params], [first_cap, cap])
        for g in groups:
            ps = [params[i] for i in g]
            total = sum((p.numel() + 7) // 8 * 8 for p in ps)    # keep every view 16-byte aligned
            flat = symm_tensor(self.comm, total, ps[0].dtype)
            flat.zero_()
            off = 0
            for p in ps:
                p.grad = flat[off:off + p.numel()].view(p.shape)
                off += (p.numel() + 7) // 8 * 8
            self.buckets.append(flat)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def zero_grad(self, set_to_none: bool = False) -> None:   # grads must stay views of the buckets
        for b in self.buckets:
            b.zero_()

    def reduce_grads(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Average the gradient buckets across the current peer group (in place, zero-copy)."""
        world = self.comm.world
        for b in self.buckets:
            self.comm.allreduce_bucket(b, b, scale=1.0 / world, post_scale=True, algo=self.algo,
                                       stream=stream)

    # ---- elastic ------------------------------------------------------------------------------
    def reform(self, new_world: int, new_rank: int, member_mask: int, epoch: int) -> None:
        self.comm.reform(new_world, new_rank, member_mask, epoch)

    @torch.no_grad()
    def sync_params(self, root: int = 0) -> None:
        """Replicate rank `root`'s parameters and buffers to the whole group through the exchange
        kernels themselves (root contributes its values, everybody else zeros)."""
        me = self.comm.rank
        tensors = [p.data for p in self.module.parameters()] + \
                  [b for b in self.module.buffers() if b.is_floating_point()]
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t)
        for dtype, ts in by_dtype.items():
            n = sum(t.numel() for t in ts)
            flat = torch.zeros((n + 7) // 8 * 8, dtype=dtype, device=ts[0].device)
            if me == root:
                torch.cat([t.reshape(-1) for t in ts], out=flat[:n])
            self.comm.allreduce_bucket(flat, flat, scale=1.0)
            off = 0
            for t in ts:
                t.copy_(flat[off:off + t.numel()].view(t.shape))
                off += t.numel()
