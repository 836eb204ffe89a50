"""What one cross-replica barrier costs, and which part of it (tok_comm_debug_barrier): K barriers
back to back in one kernel, per variant and grid size.  torchrun, one rank per GPU.  Not product."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from torch_on_k8s_b200 import _ffi  # noqa: E402
from torch_on_k8s_b200.comm import Communicator  # noqa: E402

NAMES = {0: "production", 1: "p2p_flags", 2: "signal_only_multicast", 3: "signal_only_p2p",
         4: "production_after_64KiB_of_peer_stores"}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Communicator("barbench", rank, world, local,
                        rendezvous_path="/tmp/tok8s-bar-%s" % os.environ.get("MASTER_PORT", "0"))
    st = torch.cuda.Stream()
    L = _ffi.lib()
    out = []
    K = 2000
    for ctas in (1, 16, 64):
        for variant in (0, 1, 2, 3, 4):
            if variant in (2,) and not comm.caps().multicast:
                continue
            with torch.cuda.stream(st):
                _ffi.check(L.tok_comm_debug_barrier(comm._h, variant, ctas, 200, st.cuda_stream))
                st.synchronize()
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                _ffi.check(L.tok_comm_debug_barrier(comm._h, variant, ctas, K, st.cuda_stream))
                e1.record(st)
                st.synchronize()
            comm.status()
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / K], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            row = dict(world=world, ctas=ctas, variant=NAMES[variant], us_per_barrier=round(float(t.item()), 3),
                       multicast=int(comm.caps().multicast))
            out.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    # kernel launch + exit overhead on the same stream: an empty-ish launch (1 barrier) back to back
    with torch.cuda.stream(st):
        for _ in range(20):
            _ffi.check(L.tok_comm_debug_barrier(comm._h, 0, 1, 1, st.cuda_stream))
        st.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(200):
            _ffi.check(L.tok_comm_debug_barrier(comm._h, 0, 1, 1, st.cuda_stream))
        e1.record(st)
        st.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 200], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    row = dict(world=world, what="one launch = one barrier, back to back launches", us_per_launch=round(float(t.item()), 3))
    out.append(row)
    if rank == 0:
        print(json.dumps(row), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "barrier_bench_n%d.json" % world), "w") as f:
            json.dump(out, f, indent=1)
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
