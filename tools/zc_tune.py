"""Grid-size tuning of the zero-copy exchange at the DDP bucket sizes (isolated, back to back, inputs
rotated over > L2): TOK_ZC_CTAS in {8,16,32,64} x {22.9 MB, 28.3 MB, both alternating}.  torchrun, one
rank per GPU.  Not part of the product."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from torch_on_k8s_b200.comm import Communicator  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    st = torch.cuda.Stream()
    out = []
    sizes = [22857856, 28256208]
    configs = [(int(c), int(u)) for c in os.environ.get("CTAS", "8,16,32,64").split(",")
               for u in os.environ.get("UNROLL", "8").split(",")]
    for ctas, unroll in configs:
        os.environ["TOK_ZC_CTAS"] = str(ctas)
        os.environ["TOK_NVLS_UNROLL"] = str(unroll)
        comm = Communicator("zctune", rank, world, local,
                            rendezvous_path="/tmp/tok8s-zct-%s-%d-%d" % (os.environ["MASTER_PORT"], ctas, unroll))
        sets = [[comm.symm_empty(sz // 2, torch.bfloat16).normal_() for sz in sizes] for _ in range(4)]
        for algo, name in [(0, "auto")] + ([(3, "two_shot_inplace")] if os.environ.get("TWO_SHOT") else []):
            with torch.cuda.stream(st):
                def run(i):
                    for t in sets[i % 4]:
                        comm.allreduce_bucket(t, t, scale=1.0 / world, algo=algo, stream=st)
                for i in range(5):
                    run(i)
                st.synchronize()
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for i in range(30):
                    run(i)
                e1.record(st)
                st.synchronize()
            comm.status()
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / 60], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = float(t.item())
            busbw = (sum(sizes) / 2) / (us * 1e-6) / 1e9 * 2 * (world - 1) / world
            row = dict(world=world, zc_ctas=ctas, nvls_unroll=unroll, algo=name, kernel=comm.last_algo(),
                       us_per_bucket=round(us, 2), busbw_gbs=round(busbw, 1), frac_of_900=round(busbw / 900, 3))
            out.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
        del sets
        comm.close()
        dist.barrier()
    if rank == 0:
        with open(os.path.join(ROOT, "gpurun_out", os.environ.get("OUT", "zc_tune_n%d.json" % world)), "w") as f:
            json.dump(out, f, indent=1)
        best = min((r for r in out if r["algo"] == "auto"), key=lambda r: r["us_per_bucket"])
        with open(os.path.join(ROOT, "gpurun_out", "zc_best.env"), "w") as f:
            f.write("export TOK_ZC_CTAS=%d TOK_NVLS_UNROLL=%d\n" % (best["zc_ctas"], best["nvls_unroll"]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
