"""BASELINE config 3 ("Two queued TorchJobs (ResNet-50 + BERT-base) under pkg/coordinator WRR, gang
MinMember=4 each on 8 GPUs") through the CONTROLLER, plus a third job that has to wait for a gang's
worth of free GPUs.

  python tools/run_cfg4.py [--out gpurun_out/cfg4.json]

Jobs: resnet50 (queue team-a, 1 master + 3 workers, DDP over libtok8s' comm hook), bert-base (queue
team-b, 1 + 3, ElasticDataParallel), and — submitted once both are admitted — resnet50-late (queue
team-a, 1 + 3): the box has 8 slots, so it stays queued until a gang of 4 is free.  Recorded: coordinator dequeue
order and times (WRR over the two queues, one dequeue per 100 ms tick), gang admissions, queue waits,
and each job's throughput from its own progress lines.  Measurement harness, not product.
"""
import argparse
import json
import os
import re
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torch_on_k8s_b200.controller import Controller  # noqa: E402
from torch_on_k8s_b200.netutil import free_port  # noqa: E402


def job(name, queue, cmd, workers=3):
    c = {"name": "torch", "image": "local", "command": cmd,
         "ports": [{"name": "torchjob-port", "containerPort": free_port()}]}
    return {"metadata": {"name": name, "namespace": "default"},
            "spec": {"schedulingPolicy": {"queue": queue},
                     "minMembers": {"Master": 1, "Worker": workers},
                     "torchTaskSpecs": {"Master": {"template": {"spec": {"containers": [c]}}},
                                        "Worker": {"numTasks": workers,
                                                   "template": {"spec": {"containers": [dict(c)]}}}}}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/cfg4.json")
    ap.add_argument("--gpus", type=int, default=8)
    ap.add_argument("--resnet-steps", type=int, default=150)
    ap.add_argument("--bert-steps", type=int, default=300)
    a = ap.parse_args()
    work = tempfile.mkdtemp(prefix="tok8s-cfg4-")
    logs = os.path.join(work, "logs")
    ctl = Controller(num_gpus=a.gpus, policy="wrr", log_dir=logs, rdzv_dir=work)
    py = sys.executable
    resnet = [py, os.path.join(ROOT, "workloads", "train.py"), "--model", "resnet50", "--batch", "256",
              "--steps", str(a.resnet_steps), "--log-every", "10"]
    bert = [py, os.path.join(ROOT, "workloads", "train_elastic.py"), "--model", "bert", "--batch", "16",
            "--steps", str(a.bert_steps)]
    t0 = time.time()
    uids = {"resnet50": ctl.submit(job("resnet50", "team-a", resnet)),
            "bert-base": ctl.submit(job("bert-base", "team-b", bert))}
    # the two jobs of the config first: WRR hands out one dequeue per 100 ms tick, alternating the
    # queues; both gangs fit (4 + 4 of 8 slots).  Once both are admitted a third job joins team-a's
    # queue: no gang's worth of GPUs is free, so it is held until one of the two finishes.
    t_wait = time.time()
    while time.time() - t_wait < 60 and sum(e[2] == "GangAdmitted" for e in ctl.events) < 2:
        ctl.tick()
        time.sleep(0.02)
    uids["resnet50-late"] = ctl.submit(job("resnet50-late", "team-a", resnet))
    res = ctl.run_until_done(timeout=1500)
    ev = ctl.events

    def when(uid, reason):
        xs = [t for t, u, r, m in ev if u == uid and r == reason]
        return round(xs[0] - t0, 3) if xs else None

    def resnet_rate(name):
        lats = []
        try:
            for ln in open(os.path.join(logs, name + "-master-0.log"), errors="replace"):
                m = re.match(r"Epoch: \[0\]\[\s*(\d+)/\d+\]\tTime\s+([0-9.]+)", ln)
                if m and int(m.group(1)) > 20:
                    lats.append(float(m.group(2)))
        except OSError:
            return None
        return 256 * 4 / statistics.median(lats) if lats else None

    def bert_rate(name):
        xs = []
        try:
            for ln in open(os.path.join(logs, name + "-master-0.log"), errors="replace"):
                if ln.startswith("TOK8S_STEP "):
                    r = json.loads(ln[11:])
                    if r["event"] == "step" and r["step"] > 10:
                        xs.append(r["seconds"])
        except OSError:
            return None
        return 16 * 512 * 4 / statistics.median(xs) if xs else None

    jobs = {}
    for name, uid in uids.items():
        jobs[name] = dict(result=res[uid], enqueued_s=when(uid, "JobEnqueued"),
                          dequeued_s=when(uid, "JobDequeued"), gang_admitted_s=when(uid, "GangAdmitted"),
                          finished_s=when(uid, "Job" + res[uid]),
                          queue_wait_s=(when(uid, "GangAdmitted") or 0) - (when(uid, "JobEnqueued") or 0))
    jobs["resnet50"]["images_per_s"] = resnet_rate("resnet50")
    jobs["resnet50-late"]["images_per_s"] = resnet_rate("resnet50-late")
    jobs["bert-base"]["tokens_per_s"] = bert_rate("bert-base")
    out = dict(config="BASELINE config 3: ResNet-50 + BERT-base queued under the WRR coordinator, gang "
                      "MinMember = 1 master + 3 workers each, %d GPU slots; a third 4-replica job waits "
                      "for a free gang" % a.gpus,
               dequeue_order=[u for t, u, r, m in ev if r == "JobDequeued"],
               admission_order=[u for t, u, r, m in ev if r == "GangAdmitted"],
               jobs=jobs, free_gpus_at_end=len(ctl.free_gpus))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
