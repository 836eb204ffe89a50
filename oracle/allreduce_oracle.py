"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product (torch-on-k8s_b200/); only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it, and only as a checker.

CPU restatement (numpy, IEEE fp32, round-to-nearest-even) of what one DDP gradient bucket goes
through per step in the reference-configured torchjob.  The arithmetic is NOT in the reference repo
(pure Go operator, no collective call — SURVEY.md §0.2); it lives in the un-vendored third-party
dependency PyTorch c10d (pinned here to torch 2.11.0+cu128, gloo bundled):

  * scale:      default DDP comm hook `tensor.div_(N)` then all_reduce
                (torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:18-33); the built-in
                Reducer path multiplies by the fp32 scalar 1/N before the sum,
                identical to div_ for N in {2,4,8}.
  * cast:       `bf16_compress_hook` / `_compress_hook`: `buffer.to(bf16).div_(N)` -> allreduce ->
                copy back to the bucket dtype (default_hooks.py:57-92, 116-134).
  * allreduce:  ProcessGroupGloo::allreduce = SUM over ranks.  gloo sums in the wire dtype in ring
                order; the parity target fixed by SURVEY.md §8(c) is the *fp32-accumulated* sum
                rounded once, which this oracle restates with a fixed rank order 0,1,...,N-1 so that
                it is a function (bit-reproducible) rather than a tolerance band.

Pinning: tests/test_oracle_golden.py checks this file against tests/golden/allreduce_gloo_*.npz,
which tests/golden/make_golden.py generated from live torch/gloo multi-process runs in this
container (the reference repo itself ships no tests or vectors: "parity unpinned" by the
reference, pinned by the live third-party implementation).

dtype vocabulary: "f32" -> np.float32 arrays, "bf16" -> np.uint16 bit patterns, "f16" -> np.float16.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

DTYPES = ("f32", "bf16", "f16")


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << np.uint32(16)).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16, round to nearest even (what cvt.rn.bf16.f32 / torch .to(bfloat16) do)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    lsb = (u >> np.uint32(16)) & np.uint32(1)
    rounded = (u + np.uint32(0x7FFF) + lsb) >> np.uint32(16)
    nan = np.isnan(x)
    if nan.any():
        rounded = np.where(nan, np.uint32(0x7FFF), rounded)
    return rounded.astype(np.uint16)


def to_f32(a: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "f32":
        return np.asarray(a, dtype=np.float32)
    if dtype == "bf16":
        return bf16_bits_to_f32(np.asarray(a, dtype=np.uint16))
    if dtype == "f16":
        return np.asarray(a, dtype=np.float16).astype(np.float32)
    raise ValueError(dtype)


def from_f32(x: np.ndarray, dtype: str) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    if dtype == "f32":
        return x.copy()
    if dtype == "bf16":
        return f32_to_bf16_bits(x)
    if dtype == "f16":
        with np.errstate(over="ignore"):
            return x.astype(np.float16)  # numpy rounds to nearest even
    raise ValueError(dtype)


def stage_wire(inp: np.ndarray, in_dtype: str, wire_dtype: str, pre: float) -> np.ndarray:
    """The fused bucket cast/scale on the way to the wire: cast_wire(f32(in) * pre).
    Follows `buffer.to(dtype).div_(N)` (default_hooks.py:71) with the multiply done in fp32."""
    x = to_f32(inp, in_dtype) * np.float32(pre)
    return from_f32(x, wire_dtype)


def allreduce_oracle(inputs: Sequence[np.ndarray], in_dtype: str, wire_dtype: str,
                     out_dtype: str, scale: float, post: bool = False) -> np.ndarray:
    """Result every replica must hold after tok_allreduce_bucket (include/tok8s.h):

        wire_r = cast_wire(f32(in_r) * pre)          pre  = scale (PRE) | 1 (POST)
        acc    = f32(wire_0) + f32(wire_1) + ...     fp32, rank order
        out    = cast_out(f32(cast_wire(acc * post)))  post = 1 (PRE) | scale (POST)
    """
    assert len(inputs) >= 1
    pre = 1.0 if post else scale
    pst = scale if post else 1.0
    wires: List[np.ndarray] = [stage_wire(a, in_dtype, wire_dtype, pre) for a in inputs]
    acc = to_f32(wires[0], wire_dtype).copy()
    for w in wires[1:]:
        acc = acc + to_f32(w, wire_dtype)
    acc = acc * np.float32(pst)
    res_wire = from_f32(acc, wire_dtype)
    return from_f32(to_f32(res_wire, wire_dtype), out_dtype)


def allreduce_f32_unrounded(inputs: Sequence[np.ndarray], in_dtype: str, wire_dtype: str,
                            scale: float, post: bool = False) -> np.ndarray:
    """The fp32 accumulator before the final rounding — reference for order-free comparisons (NVLS
    sums inside the switch in an unspecified order)."""
    pre = 1.0 if post else scale
    pst = scale if post else 1.0
    acc = None
    for a in inputs:
        w = to_f32(stage_wire(a, in_dtype, wire_dtype, pre), wire_dtype).astype(np.float64)
        acc = w if acc is None else acc + w
    return (acc * pst).astype(np.float64)


def ulp_distance(a: np.ndarray, b: np.ndarray, dtype: str) -> np.ndarray:
    """Distance in units-in-the-last-place between two arrays of storage dtype `dtype`."""
    def key(x):
        if dtype == "f32":
            u = np.asarray(x, dtype=np.float32).view(np.int32).astype(np.int64)
            return np.where(u < 0, np.int64(-(2 ** 31)) - u, u)
        if dtype == "bf16":
            u = np.asarray(x, dtype=np.uint16).view(np.int16).astype(np.int64)
        else:
            u = np.asarray(x, dtype=np.float16).view(np.int16).astype(np.int64)
        return np.where(u < 0, np.int64(-(2 ** 15)) - u, u)
    return np.abs(key(a) - key(b))


# ---- DistributedSampler -------------------------------------------------
def shard_indices(perm: Sequence[int], rank: int, world: int, drop_last: bool = False) -> List[int]:
    """torch/utils/data/distributed.py:107-145 after the permutation is drawn: pad (by wrapping) or
    truncate to a multiple of `world`, then take indices[rank::world]."""
    n = len(perm)
    idx = list(perm)
    if drop_last and n % world != 0:
        num = -(-(n - world) // world)  # ceil((n - world) / world)
    else:
        num = -(-n // world)
    total = num * world
    if not drop_last:
        pad = total - len(idx)
        if pad > 0:
            if pad <= len(idx):
                idx += idx[:pad]
            else:
                idx += (idx * (-(-pad // len(idx))))[:pad]
    else:
        idx = idx[:total]
    assert len(idx) == total
    out = idx[rank:total:world]
    assert len(out) == num
    return out
