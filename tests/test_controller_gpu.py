"""GPU tests that drive the CONTROLLER (not the Communicator directly): a TorchJob is submitted, the
coordinator dequeues it, the gang is admitted, replicas are started with the reference's env contract
and train on the GPU through libtok8s; a user-driven rescale 2 -> 4 -> 2 (row a7; BASELINE config 2 in
miniature) re-forms the peer group in place — survivors are never restarted, joiners receive
parameters and optimizer state, dropped replicas leave at a step boundary — and two jobs queued
under WRR with MinMember gangs share the box (BASELINE config 3 in miniature).  With one visible GPU
every slot maps to cuda:0."""
import json
import os
import sys
import time

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def manifest(name, workers, queue=None, min_members=None):
    c = {"name": "torch", "image": "local",
         "command": [sys.executable, os.path.join(HERE, "gpu_elastic_replica.py")],
         "ports": [{"name": "torchjob-port", "containerPort": free_port()}],
         "resources": {"limits": {"nvidia.com/gpu": 1}}}
    m = {"metadata": {"name": name, "namespace": "default"},
         "spec": {"torchTaskSpecs": {"Master": {"template": {"spec": {"containers": [c]}}},
                                     "Worker": {"numTasks": workers,
                                                "template": {"spec": {"containers": [dict(c)]}}}}}}
    if queue:
        m["spec"]["schedulingPolicy"] = {"queue": queue}
    if min_members:
        m["spec"]["minMembers"] = min_members
    return m


def wait_for(pred, ctl, timeout):
    t0 = time.time()
    while time.time() - t0 < timeout:
        ctl.tick()
        if pred():
            return True
        time.sleep(0.05)
    return False


def dump(ctl, log_dir):
    """events + the tail of every replica log, for assertion messages"""
    out = ["events: %r" % (ctl.events[-12:],)]
    if os.path.isdir(log_dir):
        for fn in sorted(os.listdir(log_dir)):
            try:
                out.append("---- %s ----\n%s" % (fn, "".join(open(os.path.join(log_dir, fn),
                                                                   errors="replace").readlines()[-25:])))
            except OSError:
                pass
    return "\n".join(out)


def progress(log_dir, replica):
    try:
        lines = open(os.path.join(log_dir, replica + ".log")).read().splitlines()
    except OSError:
        return 0
    for ln in reversed(lines):
        if ln.startswith("Epoch: [0]["):
            return int(ln.split("[")[2].split("/")[0])
    return 0


def test_controller_rescales_gpu_job_in_place(tok_lib, n_gpus, tmp_path, monkeypatch):
    from torch_on_k8s_b200.controller import Controller
    monkeypatch.setenv("OUT_DIR", str(tmp_path))
    monkeypatch.setenv("STEPS", "2500")
    monkeypatch.setenv("TOK_RDZV_TIMEOUT_S", "90")
    for k, v in dict(TOK_MAX_CTAS="16", TOK_STAGING_MB="16", TOK_SYMM_POOL_MB="32",
                     TOK_BARRIER_TIMEOUT_MS="120000").items():
        monkeypatch.setenv(k, v)
    logs = str(tmp_path / "logs")
    ctl = Controller(num_gpus=4, log_dir=logs, rdzv_dir=str(tmp_path), drain_grace_s=60,
                     gpu_map=[i % max(n_gpus, 1) for i in range(4)], wait_ready=True)
    uid = ctl.submit(manifest("el", workers=1))
    assert wait_for(lambda: progress(logs, "el-master-0") >= 20, ctl, 180), dump(ctl, logs)
    assert ctl.scale(uid, "Worker", 3) == 1                     # world 2 -> 4
    assert wait_for(lambda: any(e[2] == "MembershipPublished" for e in ctl.events), ctl, 120), \
        dump(ctl, logs)
    at = progress(logs, "el-master-0")
    assert wait_for(lambda: progress(logs, "el-worker-2") >= at + 30, ctl, 180), dump(ctl, logs)
    assert ctl.scale(uid, "Worker", 1) == 2                     # world 4 -> 2
    assert wait_for(lambda: sum(e[2] == "SuccessfulDeletePod" for e in ctl.events) == 2, ctl, 120), \
        dump(ctl, logs)
    res = ctl.run_until_done(timeout=400)
    assert res[uid] == "Succeeded", dump(ctl, logs)
    pods = [e[3] for e in ctl.events if e[2] == "SuccessfulCreatePod"]
    assert pods == ["el-master-0", "el-worker-0", "el-worker-1", "el-worker-2"]   # nobody restarted
    # the scale-out was announced only once both joiners were up: the survivors trained on meanwhile
    t_pub = [e[0] for e in ctl.events if e[2] == "MembershipPublished"][0]
    t_new = [e[0] for e in ctl.events if e[2] == "SuccessfulCreatePod"][-1]
    assert t_pub - t_new > 1.0
    drained = [e[3] for e in ctl.events if e[2] == "DrainingPod"]
    assert sorted(drained) == ["el-worker-1", "el-worker-2"]
    assert len(ctl.free_gpus) == 4
    docs = [json.loads(e[3]) for e in ctl.events if e[2] == "MembershipPublished"]
    assert [(d["epoch"], d["world"], d["survivor_mask"]) for d in docs] == [(1, 4, 0b0011), (2, 2, 0b0011)]
    rec = {n: json.load(open(tmp_path / (n + ".json"))) for n in pods}
    ev = {n: [(h["event"], h.get("world")) for h in r["history"] if h["event"] != "step"]
          for n, r in rec.items()}
    assert ev["el-master-0"] == [("reformed", 4), ("reformed", 2)]
    assert ev["el-worker-0"] == [("reformed", 4), ("reformed", 2)]
    assert [e[0] for e in ev["el-worker-1"]] == ["joined", "dropped"]
    assert [e[0] for e in ev["el-worker-2"]] == ["joined", "dropped"]
    # replicas hold bit-identical parameters at every common step, before, during and after the
    # 4-replica phase (momentum included: the joiners received the optimizer state)
    by_step = {}
    for n, r in rec.items():
        for h in r["history"]:
            if h["event"] == "step":
                by_step.setdefault(h["step"], {})[n] = (h["world"], h["digest"])
    worlds_seen = set()
    for step, seen in by_step.items():
        vals = list(seen.values())
        assert all(v == vals[0] for v in vals), (step, seen)
        worlds_seen.add((vals[0][0], len(vals)))
    assert (4, 4) in worlds_seen and (2, 2) in worlds_seen
    assert rec["el-master-0"]["kernel"] in ("two_shot_inplace", "nvls_inplace", "one_shot", "bcast_staged")


def test_two_queued_gpu_jobs_gang_admission_under_wrr(tok_lib, n_gpus, tmp_path, monkeypatch):
    """Two TorchJobs in different queues, MinMember = whole job (1 master + 1 worker) each, on a
    3-slot box: WRR dequeues one per tick, the first gang takes 2 slots, the second gang (needs 2,
    1 free) is held all-or-nothing until the first finishes — then runs.  Both train on the GPU."""
    from torch_on_k8s_b200.controller import Controller
    monkeypatch.setenv("OUT_DIR", str(tmp_path))
    monkeypatch.setenv("STEPS", "40")
    for k, v in dict(TOK_MAX_CTAS="16", TOK_STAGING_MB="16", TOK_SYMM_POOL_MB="32",
                     MAX_WORLD="2").items():
        monkeypatch.setenv(k, v)
    ctl = Controller(num_gpus=3, log_dir=str(tmp_path / "logs"), rdzv_dir=str(tmp_path),
                     gpu_map=[i % max(n_gpus, 1) for i in range(3)])
    a = ctl.submit(manifest("ja", 1, queue="qa", min_members={"Master": 1, "Worker": 1}))
    b = ctl.submit(manifest("jb", 1, queue="qb", min_members={"Master": 1, "Worker": 1}))
    res = ctl.run_until_done(timeout=300)
    assert res == {a: "Succeeded", b: "Succeeded"}, dump(ctl, str(tmp_path / "logs"))
    admitted = [e[1] for e in ctl.events if e[2] == "GangAdmitted"]
    assert sorted(admitted) == sorted([a, b])
    first, second = admitted
    t_first_done = [e[0] for e in ctl.events if e[1] == first and e[2] == "JobSucceeded"][0]
    t_second_admit = [e[0] for e in ctl.events if e[1] == second and e[2] == "GangAdmitted"][0]
    assert t_second_admit >= t_first_done - 1.0      # held until the first gang released its slots
    assert len(ctl.free_gpus) == 3
