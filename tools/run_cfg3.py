"""BASELINE config 2 ("BERT-base bf16 seq=512 synthetic, 4 workers, elastic rescale 4 -> 8 -> 4 mid-run")
through the CONTROLLER on an 8-GPU box.

  python tools/run_cfg3.py [--out gpurun_out/cfg3.json] [--model bert] [--gpus 8]

A TorchJob (1 master + 3 workers = 4 replicas, workloads/train_elastic.py) is submitted; when the
master has done --scale-out-at steps the job is scaled to 1 + 7 (Controller.scale: a spec update, row
a7); after --phase-steps steps at world 8 it is scaled back to 1 + 3.  Survivors re-form in place
(tok_comm_reform), joiners are announced once they are up and receive parameters + optimizer state
over tok_broadcast, dropped replicas leave at a step boundary.  Reported per phase: tokens/s
(median step time), and per re-form: how long the survivors' training was paused, against what the
reference does for the same event — restart EVERY replica with a new WORLD_SIZE
(controllers/train/elastic_scale.go:210-397) — measured here as a replica's process-start -> first
completed step time in this very run.  Not part of the product (measurement harness).
"""
import argparse
import json
import os
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torch_on_k8s_b200.controller import Controller  # noqa: E402
from torch_on_k8s_b200.netutil import free_port  # noqa: E402


def records(log_dir):
    out = []
    if not os.path.isdir(log_dir):
        return out
    for fn in os.listdir(log_dir):
        try:
            for ln in open(os.path.join(log_dir, fn), errors="replace"):
                if ln.startswith("TOK8S_STEP "):
                    out.append(json.loads(ln[len("TOK8S_STEP "):]))
        except (OSError, ValueError):
            continue
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/cfg3.json")
    ap.add_argument("--model", default="bert")
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--small", type=int, default=3, help="workers in the small phases (world = 1 + this)")
    ap.add_argument("--big", type=int, default=7, help="workers in the big phase")
    ap.add_argument("--scale-out-at", type=int, default=20)
    ap.add_argument("--phase-steps", type=int, default=60)
    ap.add_argument("--steps", type=int, default=100000, help="upper bound; the run ends by phases")
    ap.add_argument("--gpu-map", default="", help="comma list slot -> physical GPU (tests on small boxes)")
    a = ap.parse_args()

    work = tempfile.mkdtemp(prefix="tok8s-cfg3-")
    logs = os.path.join(work, "logs")
    gpu_map = [int(x) for x in a.gpu_map.split(",")] if a.gpu_map else None
    ctl = Controller(num_gpus=a.gpus, log_dir=logs, rdzv_dir=work, drain_grace_s=120, wait_ready=True,
                     gpu_map=gpu_map)
    cmd = [sys.executable, os.path.join(ROOT, "workloads", "train_elastic.py"), "--model", a.model,
           "--batch", str(a.batch), "--steps", str(a.steps)]
    c = {"name": "torch", "image": "local", "command": cmd,
         "ports": [{"name": "torchjob-port", "containerPort": free_port()}]}
    manifest = {"metadata": {"name": "bert-elastic", "namespace": "default"},
                "spec": {"torchTaskSpecs": {
                    "Master": {"template": {"spec": {"containers": [c]}}},
                    "Worker": {"numTasks": a.small, "template": {"spec": {"containers": [dict(c)]}}}}}}
    t_submit = time.time()
    uid = ctl.submit(manifest)
    timeline = []

    def master_steps(world=None):
        return [r for r in records(logs) if r["event"] == "step" and r["rank"] == 0 and
                (world is None or r["world"] == world)]

    def run_until(pred, timeout, what):
        t0 = time.time()
        while time.time() - t0 < timeout:
            ctl.tick()
            if pred():
                return
            if ctl.jobs[uid].done:
                raise SystemExit("job ended while waiting for: %s\n%s" % (what, ctl.events[-10:]))
            time.sleep(0.1)
        raise SystemExit("timeout waiting for: %s\n%s" % (what, ctl.events[-10:]))

    w_small, w_big = 1 + a.small, 1 + a.big
    run_until(lambda: len(master_steps(w_small)) >= a.scale_out_at, 600, "first steps at world %d" % w_small)
    timeline.append(dict(t=time.time(), what="scale-out requested", epoch=ctl.scale(uid, "Worker", a.big)))
    run_until(lambda: len(master_steps(w_big)) >= a.phase_steps, 900, "steps at world %d" % w_big)
    n_small_before = len(master_steps(w_small))
    timeline.append(dict(t=time.time(), what="scale-in requested", epoch=ctl.scale(uid, "Worker", a.small)))
    run_until(lambda: len(master_steps(w_small)) >= n_small_before + a.phase_steps, 900,
              "steps back at world %d" % w_small)
    # done: stop the job (cleanPodPolicy Running semantics): the replicas would run to --steps
    for reps in ctl.jobs[uid].replicas.values():
        for r in reps.values():
            ctl._kill(r)
    recs = records(logs)

    def phase(world, lo_t=None, hi_t=None):
        xs = [r for r in recs if r["event"] == "step" and r["rank"] == 0 and r["world"] == world and
              (lo_t is None or r["t"] >= lo_t) and (hi_t is None or r["t"] <= hi_t)]
        if len(xs) < 4:
            return None
        sec = statistics.median(r["seconds"] for r in xs[2:])
        return dict(world=world, steps=len(xs), median_step_s=sec,
                    units_per_s=xs[0]["units"] * world / sec)
    t_out = [e for e in recs if e["event"] == "reformed" and e["world"] == w_big]
    t_in = [e for e in recs if e["event"] == "reformed" and e["world"] == w_small]
    joined = [e for e in recs if e["event"] == "joined"]
    created = {e[3]: e[0] for e in ctl.events if e[2] == "SuccessfulCreatePod"}
    first_step = {}
    for r in recs:
        if r["event"] == "step":
            first_step.setdefault(r["replica"], r["t"])
    cold = {n: first_step[n] - t for n, t in created.items() if n in first_step}
    initial = [v for n, v in cold.items() if created[n] < timeline[0]["t"]]
    t_split = t_out[0]["t"] if t_out else None
    t_split2 = t_in[0]["t"] if t_in else None
    out = dict(
        config="BASELINE config 2: %s bf16 seq 512 batch %d/replica, 1 master + %d workers -> 1 + %d "
               "-> 1 + %d, rescaled by the controller mid-run" % (a.model, a.batch, a.small, a.big, a.small),
        phases=[phase(w_small, None, t_split), phase(w_big, t_split, t_split2), phase(w_small, t_split2, None)],
        unit="tokens/s" if a.model == "bert" else "samples/s",
        reform_scale_out=dict(
            survivors_paused_s=max((e["reform_visible_s"] for e in t_out), default=None),
            of_which_state_handover_s=max((e["sync_s"] for e in t_out), default=None),
            joiner_process_start_to_joined_s=max((e["startup_s"] for e in joined), default=None),
            note="joiners start while the survivors keep training; the membership is announced once "
                 "they are up"),
        reform_scale_in=dict(
            survivors_paused_s=max((e["reform_visible_s"] for e in t_in), default=None),
            of_which_state_handover_s=max((e["sync_s"] for e in t_in), default=None)),
        restart_baseline=dict(
            per_replica_process_start_to_first_step_s=(max(initial) if initial else None),
            note="what the reference's rescale costs at least: every replica restarted with the new "
                 "WORLD_SIZE (elastic_scale.go:303-397) = process start + CUDA init + model build + "
                 "rendezvous + first step, measured on this run's own initial start; checkpoint "
                 "save/load not included"),
        events=[(round(t - t_submit, 3), r, m[:200]) for t, u, r, m in ctl.events
                if r in ("Scale", "MembershipPublished", "SuccessfulCreatePod", "DrainingPod",
                         "SuccessfulDeletePod", "GangAdmitted", "JobDequeued")],
        timeline=[dict(t=round(x["t"] - t_submit, 3), what=x["what"], epoch=x["epoch"]) for x in timeline],
    )
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: out[k] for k in ("phases", "reform_scale_out", "reform_scale_in",
                                          "restart_baseline")}, indent=1))


if __name__ == "__main__":
    main()
