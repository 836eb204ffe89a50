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


@torch.no_grad()
def broadcast_coalesced(comm: Communicator, tensors, root: int = 0, use_pool: bool = True) -> None:
    """Replicate `root`'s `tensors` (one dtype, same device) into every replica's, bit for bit: packed
    into one flat buffer, one tok_broadcast, unpacked in place on the receivers.  The flat buffer is
    taken from the symmetric pool when it fits — the root then feeds all receivers with ONE
    multimem.st stream through the NVSwitch — and released again (same order on every replica)."""
    n = sum(t.numel() for t in tensors)
    if n == 0:
        return
    dtype, dev = tensors[0].dtype, tensors[0].device
    esz = tensors[0].element_size()
    n_pad = (n * esz + 15) // 16 * 16 // esz
    ptr = C.c_void_p()
    pooled = use_pool and lib().tok_comm_symm_alloc(comm._h, n_pad * esz, C.byref(ptr)) == 0
    if pooled:
        raw = torch.as_tensor(_RawCuda(ptr.value, n_pad * esz), device=dev)
        flat = raw.view(dtype)
    else:
        flat = torch.empty(n_pad, dtype=dtype, device=dev)
    if comm.rank == root:
        torch.cat([t.reshape(-1) for t in tensors], out=flat[:n])
    comm.broadcast(flat, root)
    if comm.rank != root:
        off = 0
        for t in tensors:
            t.copy_(flat[off:off + t.numel()].view(t.shape))
            off += t.numel()
    if pooled:
        torch.cuda.current_stream(dev).synchronize()   # the segment is recycled: nothing in flight
        del flat, raw
        check(lib().tok_comm_symm_free(comm._h, ptr, n_pad * esz))


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
                 first_bucket_mb: int = 1, algo: int = 0, pool_broadcast: bool = True):
        super().__init__()
        self.module = module
        self.comm = comm
        self.algo = algo
        self.pool_broadcast = pool_broadcast   # state hand-over through pool buffers (zero-copy path)
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
        for b in self.buckets:   # PRE scale: the same function as built-in DDP at every world size
            self.comm.allreduce_bucket(b, b, scale=1.0 / world, algo=self.algo, stream=stream)

    # ---- elastic ------------------------------------------------------------------------------
    def reform(self, new_world: int, new_rank: int, member_mask: int, epoch: int) -> None:
        self.comm.reform(new_world, new_rank, member_mask, epoch)

    @torch.no_grad()
    def _broadcast_tensors(self, tensors, root: int) -> None:
        """Coalesced broadcast (dist._broadcast_coalesced's job, torch/nn/parallel/distributed.py:1032)
        through tok_broadcast: tensors are packed per dtype into one flat buffer, replicated
        bit-for-bit from `root`, and unpacked in place on the receivers."""
        me = self.comm.rank
        by_dtype = {}
        for t in tensors:
            by_dtype.setdefault(t.dtype, []).append(t)
        for dtype, ts in by_dtype.items():
            n = sum(t.numel() for t in ts)
            if n == 0:
                continue
            broadcast_coalesced(self.comm, ts, root, use_pool=self.pool_broadcast)

    @torch.no_grad()
    def sync_params(self, root: int = 0) -> None:
        """Replicate rank `root`'s parameters and ALL buffers (BatchNorm's integer
        num_batches_tracked included) to the whole group — what a replica that joins at an elastic
        re-form needs before its first step."""
        self._broadcast_tensors([p.data for p in self.module.parameters()] +
                                [b for b in self.module.buffers()], root)

    @torch.no_grad()
    def sync_optimizer_state(self, optimizer: torch.optim.Optimizer, root: int = 0) -> None:
        """Replicate `root`'s optimizer state (SGD momentum buffers, Adam moments and step counts).
        Without it a joiner would apply the same averaged gradients with different momentum and the
        replicas would drift apart for good.  State is created lazily by torch optimizers, so which
        entries exist is itself part of what must be replicated: the root announces its layout (one
        small broadcast), everybody allocates missing entries, then the tensors follow."""
        me = self.comm.rank
        dev = next(self.module.parameters()).device
        params = [p for g in optimizer.param_groups for p in g["params"]]
        # layout = for every parameter, which tensor-valued state keys the root holds (bitmask over
        # the sorted union of known keys; python scalars such as `step` ints travel as f64)
        keys = ["momentum_buffer", "exp_avg", "exp_avg_sq", "max_exp_avg_sq", "step", "square_avg",
                "acc_delta", "sum"]
        layout = torch.zeros(len(params), dtype=torch.int64, device=dev)
        if me == root:
            for i, p in enumerate(params):
                st = optimizer.state.get(p, {})
                layout[i] = sum(1 << k for k, name in enumerate(keys) if st.get(name) is not None)
        self.comm.broadcast(layout, root)
        masks = layout.tolist()
        tensors, scalars = [], []
        for p, m in zip(params, masks):
            st = optimizer.state[p] if m else optimizer.state.get(p, {})
            for k, name in enumerate(keys):
                if not (m >> k) & 1:
                    if name in st and me != root:
                        del st[name]
                    continue
                v = st.get(name)
                if torch.is_tensor(v):
                    if v.device != p.device and v.dim() == 0:   # host-side step counter
                        scalars.append((st, name, v))
                    else:
                        tensors.append(v)
                elif v is None:          # joiner: allocate what the root has
                    if name == "step":
                        st[name] = torch.zeros((), dtype=torch.float32)
                        scalars.append((st, name, st[name]))
                    else:
                        st[name] = torch.zeros_like(p, memory_format=torch.preserve_format)
                        tensors.append(st[name])
                else:                    # python number
                    scalars.append((st, name, v))
        self._broadcast_tensors(tensors, root)
        if scalars:
            flat = torch.tensor([float(v) for _, _, v in scalars], dtype=torch.float64, device=dev)
            self.comm.broadcast(flat, root)
            for (st, name, old), v in zip(scalars, flat.tolist()):
                if torch.is_tensor(old):
                    old.fill_(v)
                else:
                    st[name] = type(old)(v)
