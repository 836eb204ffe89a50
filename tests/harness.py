"""Multi-replica test harness: runs `world` replicas as processes (the product shape: one process per
replica) or as threads of one process, on one GPU (all replicas share cuda:0 — what the round-end
`pytest -m gpu` box offers) or on one GPU per replica.

Every replica regenerates the seeded inputs of ALL replicas on the CPU, pushes its own to the GPU,
calls the CUDA path through the C ABI (torch_on_k8s_b200.comm.Communicator -> libtok8s) and compares
the result bit-for-bit with oracle/allreduce_oracle.py.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import sys
import tempfile
import threading
import time
import traceback
from typing import Dict, List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NP2TOK = {"f32": 0, "bf16": 1, "f16": 2}


def torch_dtype(name):
    import torch
    return {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[name]


def gen_input(seed: int, rank: int, count: int, dtype: str, pattern: str = "randn") -> np.ndarray:
    """Seeded synthetic bucket of replica `rank` in storage form (see oracle dtype vocabulary)."""
    from oracle.allreduce_oracle import from_f32
    rs = np.random.RandomState((seed * 1000003 + rank * 7919 + 17) % (2 ** 31 - 1))
    if pattern == "randn":
        x = rs.standard_normal(count).astype(np.float32)
    elif pattern == "ints":  # exactly representable: sums are order independent
        x = (rs.randint(-8, 9, size=count)).astype(np.float32) * np.float32(0.25)
    elif pattern == "rank":  # rank+1 everywhere (nccl-tests style)
        x = np.full(count, float(rank + 1), dtype=np.float32)
    elif pattern == "wide":  # wide dynamic range: exposes accumulation-order differences
        x = (rs.standard_normal(count) * np.exp(rs.uniform(-12, 12, size=count))).astype(np.float32)
    else:
        raise ValueError(pattern)
    return from_f32(x, dtype)


def to_torch(a: np.ndarray, dtype: str, device):
    import torch
    if dtype == "bf16":
        t = torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16)
    elif dtype == "f16":
        t = torch.from_numpy(a.copy())
    else:
        t = torch.from_numpy(a.copy())
    return t.to(device)


def from_torch(t, dtype: str) -> np.ndarray:
    import torch
    t = t.detach().cpu().contiguous()
    if dtype == "bf16":
        return t.view(torch.int16).numpy().view(np.uint16)
    return t.numpy()


def bits_equal(a: np.ndarray, b: np.ndarray, dtype: str) -> bool:
    """Bit-exact comparison; NaNs compare equal to NaNs whatever their payload (inf - inf in a
    reduced-range wire dtype: numpy and cvt.rn produce different quiet-NaN encodings)."""
    from oracle.allreduce_oracle import to_f32
    if dtype == "f32":
        ua, ub = (np.asarray(x, np.float32).view(np.uint32) for x in (a, b))
    elif dtype == "f16":
        ua, ub = (np.asarray(x, np.float16).view(np.uint16) for x in (a, b))
    else:
        ua, ub = (np.asarray(x, np.uint16) for x in (a, b))
    same = ua == ub
    if same.all():
        return True
    both_nan = np.isnan(to_f32(a, dtype)) & np.isnan(to_f32(b, dtype))
    return bool((same | both_nan).all())


def run_cases(rank: int, world: int, device: int, path: str, cases: List[dict],
              job: str = "harness", use_mempool: bool = True) -> List[dict]:
    """Body of one replica.  Returns one result dict per case.  use_mempool: pool buckets come from a
    torch.cuda.MemPool over tok_pool_malloc (one communicator per process) — replicas that are
    threads of one process allocate from their pool directly instead."""
    import torch
    from oracle.allreduce_oracle import allreduce_f32_unrounded, allreduce_oracle, to_f32, ulp_distance
    from torch_on_k8s_b200.comm import Communicator

    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    comm = Communicator(job, rank, world, device, rendezvous_path=path)
    stream = torch.cuda.Stream(device=dev)
    results = []
    if use_mempool:
        symm_empty = comm.symm_empty
    else:
        from torch_on_k8s_b200.elastic_dp import symm_tensor

        def symm_empty(n, dt):
            return symm_tensor(comm, n, dt)
    comm._test_symm_empty = symm_empty
    caps = comm.caps()
    try:
        for ci, case in enumerate(cases):
            count = case["count"]
            din, dw, dout = case["in"], case["wire"], case["out"]
            scale = case.get("scale", 1.0 / world)
            post = case.get("post", False)
            algo = case.get("algo", 0)
            pattern = case.get("pattern", "randn")
            seed = case.get("seed", 1234 + ci)
            inplace = case.get("inplace", din == dout)
            symm = case.get("symm", False)        # bucket allocated in the symmetric pool: zero-copy
            if algo == 4 and not caps.multicast:
                results.append(dict(case=case, skipped="no multicast"))
                continue
            if algo == 1 and world != 1:
                continue
            if case.get("bcast") is not None:
                results.append(_run_bcast(comm, case, rank, world, dev, stream, ci))
                continue
            if case.get("golden"):
                results.append(_run_golden(comm, case, rank, world, dev, stream))
                continue
            inputs = [gen_input(seed, r, count, din, pattern) for r in range(world)]
            want = allreduce_oracle(inputs, din, dw, dout, scale, post)
            with torch.cuda.stream(stream):
                if symm:
                    if case.get("skew") and rank == world - 1:
                        _extra = symm_empty(count, torch_dtype(din))  # steals x's block
                    x = symm_empty(count, torch_dtype(din))
                    x.copy_(to_torch(inputs[rank], din, dev))
                    assert comm.in_symmetric_pool(x)
                    y = x
                else:
                    x = to_torch(inputs[rank], din, dev)
                    y = x if inplace else torch.empty(count, dtype=torch_dtype(dout), device=dev)
                t0 = time.perf_counter()
                arrived = False
                if case.get("split"):   # the two-call form: arrival kernel, then the exchange
                    arrived = comm.bucket_arrive(x, scale=scale, post_scale=post, algo=algo,
                                                 stream=stream)
                comm.allreduce_bucket(x, y, scale=scale, wire_dtype=torch_dtype(dw),
                                      post_scale=post, algo=algo, arrived=arrived,
                                      elide=not case.get("no_elide", False), stream=stream)
                stream.synchronize()
                ms = (time.perf_counter() - t0) * 1e3
            if case.get("skew"):
                try:
                    comm.status()
                    results.append(dict(case=case, exact=False, rank=rank, ms=ms, max_ulp=-1,
                                        mismatch=-1, note="asymmetric bucket was not detected"))
                except Exception as e:  # noqa: BLE001
                    results.append(dict(case=case, exact="same symmetric-pool offset" in str(e),
                                        rank=rank, ms=ms, max_ulp=0, mismatch=0, normwise=0.0))
                break
            comm.status()
            got = from_torch(y, dout)
            exact = bits_equal(got, want, dout)
            res = dict(case=case, exact=bool(exact), ms=ms, rank=rank, kernel=comm.last_algo())
            if case.get("expect_kernel") and res["kernel"] != case["expect_kernel"] and \
                    not (case["expect_kernel"] == "nvls_inplace" and not caps.multicast):
                res["exact"] = False
                res["note"] = "took %s, expected %s" % (res["kernel"], case["expect_kernel"])
                res.update(max_ulp=-1, mismatch=-1, normwise=-1.0)
                results.append(res)
                continue
            if not exact:
                ulp = ulp_distance(got, want, dout)
                ulp = np.where(np.isnan(to_f32(got, dout)) & np.isnan(to_f32(want, dout)), 0, ulp)
                res["max_ulp"] = int(ulp.max())
                res["mismatch"] = int((ulp != 0).sum())
                ref = allreduce_f32_unrounded(inputs, din, dw, scale, post)
                err = np.abs(to_f32(got, dout).astype(np.float64) - ref)
                denom = max(float(np.abs(ref).max()), 1e-30)
                res["normwise"] = float(err.max() / denom)
                bad = np.nonzero(ulp)[0][:4]
                res["first_bad"] = [int(i) for i in bad]
            results.append(res)
        res_launch = comm.launches()
        results.append(dict(launches=res_launch, multicast=int(caps.multicast), rank=rank))
    finally:
        comm.close()
    return results


def _run_bcast(comm, case, rank, world, dev, stream, ci):
    """tok_broadcast: every replica starts from its own seeded bytes, ends with the root's."""
    import torch
    root, nbytes = case["bcast"], case["count"]
    src = [np.random.RandomState(7000 + 13 * ci + r).randint(0, 256, size=nbytes, dtype=np.uint8)
           for r in range(world)]
    with torch.cuda.stream(stream):
        if case.get("symm"):
            t = comm._test_symm_empty(nbytes, torch.uint8)
        else:
            t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        t.copy_(torch.from_numpy(src[rank]).to(dev))
        comm.broadcast(t, root, stream=stream)
        stream.synchronize()
    comm.status()
    got = t.cpu().numpy()
    exact = bool(np.array_equal(got, src[root]))
    return dict(case=case, exact=exact, rank=rank, ms=0.0, kernel=comm.last_algo(),
                max_ulp=0 if exact else -1, mismatch=int((got != src[root]).sum()), normwise=0.0)


def _run_golden(comm, case, rank, world, dev, stream):
    """The committed gloo fixtures (tests/golden/allreduce_gloo_n*.npz) through the ZERO-COPY path DDP's
    buckets take: pool buckets, arrival + in-place exchange, 1/N PRE.  P2P in-place kernel: the fp32
    result within 1e-5 norm-wise of gloo's (bit-exact at N=2), bf16 within 1 ulp; NVLS in place: the
    same tolerances (in-switch summation order is unspecified)."""
    import torch
    from oracle import allreduce_oracle as O
    with np.load(case["golden"]) as z:
        x32 = z["randn_x32_r%d" % rank]
        xb = z["randn_xb_r%d" % rank]
        want32 = z["randn_f32_prescaled"]
        wantb = O.f32_to_bf16_bits(z["randn_bf16in_f32_prescaled"])
    n32, nb = (x32.size // 4) * 4, (xb.size // 8) * 8
    with torch.cuda.stream(stream):
        zx = comm._test_symm_empty(n32, torch.float32)
        zx.copy_(torch.from_numpy(x32[:n32].copy()).to(dev))
        zb = comm._test_symm_empty(nb, torch.bfloat16)
        zb.copy_(to_torch(xb[:nb], "bf16", dev))
        comm.allreduce_bucket(zx, zx, scale=1.0 / world, algo=case.get("algo", 3), stream=stream)
        kernel = comm.last_algo()
        comm.allreduce_bucket(zb, zb, scale=1.0 / world, algo=case.get("algo", 3), stream=stream)
        stream.synchronize()
    comm.status()
    got32, gotb = zx.cpu().numpy(), from_torch(zb, "bf16")
    err = float(np.abs(got32.astype(np.float64) - want32[:n32].astype(np.float64)).max() /
                np.abs(want32).max())
    ulp = O.ulp_distance(gotb, wantb[:nb], "bf16")
    ok = kernel in ("two_shot_inplace", "nvls_inplace") and err <= 1e-5 and int(ulp.max()) <= 1
    if world == 2 and kernel == "two_shot_inplace":
        ok = ok and bool(np.array_equal(got32.view(np.uint32), want32[:n32].view(np.uint32)))
    return dict(case=case, exact=bool(ok), rank=rank, ms=0.0, kernel=kernel, normwise=err,
                max_ulp=int(ulp.max()), mismatch=int((ulp != 0).sum()))


def _proc_entry(rank, world, device, path, cases, job, env, q):
    try:
        os.environ.update(env or {})
        q.put((rank, "ok", run_cases(rank, world, device, path, cases, job)))
    except Exception:  # noqa: BLE001
        q.put((rank, "error", traceback.format_exc()))


def launch(world: int, cases: List[dict], *, devices: Optional[List[int]] = None,
           mode: str = "proc", timeout: float = 300.0, env: Optional[Dict[str, str]] = None,
           job: str = "harness") -> Dict[int, list]:
    """Run `cases` on `world` replicas; returns {rank: results}.  Raises on replica failure."""
    devices = devices or [0] * world
    tmp = tempfile.mkdtemp(prefix="tok8s-")
    path = os.path.join(tmp, "rdzv")
    out: Dict[int, list] = {}
    if mode == "proc":
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_proc_entry,
                             args=(r, world, devices[r], path, cases, job, env, q), daemon=True)
                 for r in range(world)]
        for p in procs:
            p.start()
        deadline = time.time() + timeout
        errors = []
        try:
            while len(out) + len(errors) < world:
                left = deadline - time.time()
                if left <= 0:
                    raise TimeoutError("replicas did not finish within %.0f s" % timeout)
                try:
                    rank, status, payload = q.get(timeout=min(left, 5.0))
                except Exception:  # queue.Empty
                    if any(p.exitcode not in (None, 0) for p in procs):
                        dead = [i for i, p in enumerate(procs) if p.exitcode not in (None, 0)]
                        raise RuntimeError("replica(s) %s died (exit codes %s)" %
                                           (dead, [procs[i].exitcode for i in dead]))
                    continue
                if status == "ok":
                    out[rank] = payload
                else:
                    errors.append((rank, payload))
        finally:
            for p in procs:
                p.join(timeout=10)
                if p.is_alive():
                    p.kill()
        if errors:
            raise RuntimeError("replica %d failed:\n%s" % errors[0])
    elif mode == "thread":
        if env:
            os.environ.update(env)
        errs = []

        def body(r):
            try:
                out[r] = run_cases(r, world, devices[r], path, cases, job, use_mempool=False)
            except Exception:  # noqa: BLE001
                errs.append((r, traceback.format_exc()))

        ts = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout)
        if any(t.is_alive() for t in ts):
            raise TimeoutError("replica threads did not finish within %.0f s" % timeout)
        if errs:
            raise RuntimeError("replica %d failed:\n%s" % errs[0])
    else:
        raise ValueError(mode)
    return out


def summarize(results: Dict[int, list]) -> dict:
    total = bad = skipped = 0
    worst = []
    for rank, rs in results.items():
        for r in rs:
            if "case" not in r:
                continue
            if "skipped" in r:
                skipped += 1
                continue
            total += 1
            if not r["exact"]:
                bad += 1
                worst.append(r)
    return dict(total=total, bad=bad, skipped=skipped, worst=worst[:8])


TRIPLES_CORE = [("f32", "f32", "f32"), ("bf16", "bf16", "bf16"), ("f16", "f16", "f16"),
                ("f32", "bf16", "f32")]
TRIPLES_MIXED = [("bf16", "f32", "bf16"), ("f32", "bf16", "bf16"), ("f32", "f16", "f32"),
                 ("bf16", "bf16", "f32"), ("f16", "f32", "f32")]


def standard_cases(world: int, algos, quick: bool = False) -> List[dict]:
    """The parity matrix: dtype triples x ragged/edge sizes x PRE/POST scale x algorithms."""
    counts = [1, 7, 8, 9, 1000, 4097, 65536 + 3, (1 << 20) + 5]
    if quick:
        counts = [1, 9, 4097, 65536 + 3]
    cases = []
    seed = 100
    for algo in algos:
        for (a, w, o) in TRIPLES_CORE:
            for n in counts:
                seed += 1
                cases.append(dict(count=n, **{"in": a, "wire": w, "out": o}, algo=algo, seed=seed,
                                  scale=1.0 / world))
        for (a, w, o) in TRIPLES_MIXED:
            for n in (9, 4097, 65536 + 3):
                seed += 1
                cases.append(dict(count=n, **{"in": a, "wire": w, "out": o}, algo=algo, seed=seed,
                                  scale=1.0 / world))
        # POST scale with a non power-of-two factor, wide dynamic range, out-of-place
        for (a, w, o) in TRIPLES_CORE:
            seed += 1
            cases.append(dict(count=30011, **{"in": a, "wire": w, "out": o}, algo=algo, seed=seed,
                              scale=1.0 / 3.0, post=True, pattern="wide", inplace=False))
            seed += 1
            cases.append(dict(count=30011, **{"in": a, "wire": w, "out": o}, algo=algo, seed=seed,
                              scale=1.0 / 3.0, post=False, pattern="wide", inplace=False))
    return cases
