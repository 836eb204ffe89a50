"""Host logic behind the C ABI (csrc/ctl_*.cpp) against the pure-Python oracle restatement of the Go
reference (oracle/controlplane_oracle.py) and against hand-derived known answers from SURVEY.md
Appendix A.  CPU only — the reference ships no tests (parity unpinned by the reference)."""
import copy
import json
import random

import pytest

from oracle import controlplane_oracle as O
from torch_on_k8s_b200 import _ffi
from torch_on_k8s_b200.coordinator import Coordinator
from torch_on_k8s_b200.elastic import ElasticPolicy, parse_log_line
from torch_on_k8s_b200.job import TorchJob, set_feature_gates, should_failover

NOW = "2026-01-01T00:00:00Z"


def manifest(name="job", workers=3, master_key="Master", worker_key="Worker", aimaster=False,
             queue=None, priority=None, min_members=None, elastic=False, torchelastic=None,
             namespace="ns", gpus=None, spot=0):
    def tmpl():
        c = {"name": "torch", "image": "img"}
        if gpus is not None:
            c["resources"] = {"limits": {"nvidia.com/gpu": gpus}}
        return {"spec": {"containers": [c, {"name": "sidecar", "image": "s"}]}}
    specs = {master_key: {"template": tmpl()}}
    if workers is not None:
        specs[worker_key] = {"numTasks": workers, "template": tmpl()}
        if spot:
            specs[worker_key]["spotTaskSpec"] = {"numSpotTasks": spot, "labels": {"spot": "1"}}
    if aimaster:
        specs["AIMaster"] = {"numTasks": 1, "template": tmpl()}
    m = {"metadata": {"name": name, "namespace": namespace, "generation": 3},
         "spec": {"torchTaskSpecs": specs}}
    sp = {}
    if queue:
        sp["queue"] = queue
    if priority is not None:
        sp["priority"] = priority
    if sp:
        m["spec"]["schedulingPolicy"] = sp
    if min_members:
        m["spec"]["minMembers"] = min_members
    if elastic:
        m["metadata"]["annotations"] = {"distributed.io/enable-elastic-training": "true"}
    if torchelastic:
        m["spec"]["enableTorchElastic"] = True
        m["spec"]["torchElasticPolicy"] = torchelastic
    return m


@pytest.fixture(autouse=True)
def default_gates(tok_lib):
    tok_lib.tok_set_feature_gates(_ffi.TOK_GATES_DEFAULT)
    yield
    tok_lib.tok_set_feature_gates(_ffi.TOK_GATES_DEFAULT)


# ---- defaults (Appendix A.1) ---------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(), dict(master_key="mAsTeR", worker_key="worker"),
                                dict(aimaster=True), dict(workers=None),
                                dict(min_members={"Master": 1, "Worker": 2})])
def test_defaults_match_oracle(kw):
    m = manifest(**kw)
    got = TorchJob(copy.deepcopy(m)).to_dict()
    want, _ = O.set_defaults(m)
    assert got == want


def test_defaults_known_answers():
    d = TorchJob(manifest(master_key="master", worker_key="WORKER", workers=None)).to_dict()
    assert d["apiVersion"] == "train.distributed.io/v1alpha1" and d["kind"] == "TorchJob"
    assert d["spec"]["clenPodPolicy"] == "None"              # wire name keeps the reference's typo
    master = d["spec"]["torchTaskSpecs"]["Master"]
    assert master["numTasks"] == 1 and master["restartPolicy"] == "ExitCode"
    torch_c = master["template"]["spec"]["containers"][0]
    assert torch_c["ports"] == [{"name": "torchjob-port", "containerPort": 23456}]
    assert all(c["terminationMessagePolicy"] == "FallbackToLogsOnError"
               for c in master["template"]["spec"]["containers"])
    assert "ports" not in master["template"]["spec"]["containers"][1]
    assert d["spec"]["minMembers"] == {"Master": 1}
    d = TorchJob(manifest(workers=None, worker_key="Worker")).to_dict()
    w = TorchJob({"metadata": {"name": "x"}, "spec": {"torchTaskSpecs": {
        "Master": {}, "worker": {}}}}).to_dict()["spec"]["torchTaskSpecs"]["Worker"]
    assert w["numTasks"] == 1 and w["restartPolicy"] == "OnFailure"


def test_parse_errors():
    for bad in ("{", "[]", json.dumps({"metadata": {}, "spec": {"torchTaskSpecs": {}}}),
                json.dumps({"metadata": {"name": "a"}, "spec": {}}),
                json.dumps({"metadata": {"name": "a"}, "kind": "Pod", "spec": {"torchTaskSpecs": {}}}),
                json.dumps({"metadata": {"name": "a"}, "spec": {"torchTaskSpecs": {"Master": {"numTasks": -1}}}})):
        with pytest.raises(_ffi.TokError) as e:
            TorchJob(bad)
        assert e.value.code == _ffi.TOK_ERR_INVALID


def test_unknown_fields_round_trip():
    m = manifest()
    m["spec"]["torchTaskSpecs"]["Master"]["template"]["spec"]["nodeSelector"] = {"k": "vé\"\\"}
    m["metadata"]["labels"] = {"a": "b"}
    d = TorchJob(m).to_dict()
    assert d["spec"]["torchTaskSpecs"]["Master"]["template"]["spec"]["nodeSelector"] == {"k": "vé\"\\"}
    assert d["metadata"]["labels"] == {"a": "b"}


# ---- replica identity (Appendix A.2) ----------------------------------------------------------------------
@pytest.mark.parametrize("workers", range(0, 8))
def test_rank_table(workers):
    j = TorchJob(manifest(name="rn50", workers=workers if workers else None))
    world = 1 + workers
    m = j.cluster_spec("master", 0)
    assert m["name"] == "rn50-master-0" and m["rank"] == 0 and m["worldSize"] == world
    env = {e["name"]: e["value"] for e in m["env"]}
    assert env == {"MASTER_PORT": "23456", "MASTER_ADDR": "localhost", "RANK": "0",
                   "PYTHONUNBUFFERED": "0", "WORLD_SIZE": str(world)}
    assert [e["name"] for e in m["env"]] == ["MASTER_PORT", "MASTER_ADDR", "RANK",
                                             "PYTHONUNBUFFERED", "WORLD_SIZE"]
    assert m["labels"]["task-role"] == "master" and m["restartPolicy"] == "Never"
    for i in range(workers):
        w = j.cluster_spec("worker", i)
        env = {e["name"]: e["value"] for e in w["env"]}
        assert w["name"] == "rn50-worker-%d" % i and w["rank"] == i + 1
        assert env["RANK"] == str(i + 1) and env["MASTER_ADDR"] == "rn50-master-0"
        assert env["WORLD_SIZE"] == str(world) and w["restartPolicy"] == "OnFailure"
        assert w["labels"] == {"group-name": "train.distributed.io", "job-name": "rn50",
                               "task-type": "worker", "task-index": str(i)}
        assert w["annotations"]["scheduling.k8s.io/group-name"] == "rn50-worker"


def test_cluster_spec_matches_oracle_matrix():
    for kw in (dict(), dict(elastic=True), dict(aimaster=True),
               dict(torchelastic={"rendezvousBackend": "etcd", "rendezvousEndpoint": "h:2379",
                                  "numMinReplicas": 2, "numMaxReplicas": 8}),
               dict(torchelastic={"rendezvousBackend": "c10d", "rendezvousEndpoint": "e"}),
               dict(name="a/b")):
        m = manifest(**kw)
        j = TorchJob(copy.deepcopy(m))
        d, _ = O.set_defaults(m)
        pairs = [("master", 0)] + [("worker", i) for i in range(3)] + ([("aimaster", 0)] if kw.get("aimaster") else [])
        for tt, i in pairs:
            got = j.cluster_spec(tt, i)
            want = O.cluster_spec(d, tt, i)
            assert got["name"] == want["name"] and got["rank"] == want["rank"]
            assert [(e["name"], e["value"]) for e in got["env"]] == want["env"]
            assert got["args"] == want["args"] and got["labels"] == want["labels"]
            assert got["annotations"] == want["annotations"]
            assert got["restartPolicy"] == want["restartPolicy"]
            assert got["initContainers"] == want["initContainers"]
            assert got.get("finalizers", []) == want["finalizers"]


def test_cluster_spec_errors_and_gates():
    j = TorchJob(manifest())
    with pytest.raises(_ffi.TokError) as e:
        j.cluster_spec("master", 1)
    assert "only a single master with index=0" in str(e.value)
    with pytest.raises(_ffi.TokError):
        j.cluster_spec("chief", 0)
    nomaster = TorchJob({"metadata": {"name": "w"}, "spec": {"torchTaskSpecs": {"Worker": {}}}})
    with pytest.raises(_ffi.TokError) as e:
        nomaster.cluster_spec("worker", 0)
    assert "port" in str(e.value)
    set_feature_gates(TorchLocalMasterAddr=False, GangScheduling=False)
    m = j.cluster_spec("master", 0)
    assert {e["name"]: e["value"] for e in m["env"]}["MASTER_ADDR"] == "job-master-0"
    assert "scheduling.k8s.io/group-name" not in m["annotations"]


def test_torchelastic_args_known_answer():
    j = TorchJob(manifest(name="el", workers=4, torchelastic={
        "rendezvousBackend": "etcd", "rendezvousEndpoint": "etcd:2379", "numMinReplicas": 4,
        "numMaxReplicas": 8, "numWorkersPerNodePolicy": 1}))
    assert j.cluster_spec("worker", 0)["args"] == [
        "--rdzv_backend=etcd", "--rdzv_endpoint=etcd:2379", "--rdzv_id=el", "--nproc_per_node=1",
        "--nnodes=4:8"]
    # min/max default to the WORKER count (intended; the reference reads the master's, §2.3)
    j = TorchJob(manifest(name="el", workers=4, torchelastic={"rendezvousBackend": "c10d",
                                                               "rendezvousEndpoint": "x"}))
    assert j.cluster_spec("master", 0)["args"][-1] == "--nnodes=4:4"


# ---- DAG / gang (Appendix A.4) ---------------------------------------------------------------------------------
def test_dag_gate():
    m = manifest(aimaster=True)
    j = TorchJob(copy.deepcopy(m))
    d, dep = O.set_defaults(m)
    cases = [{}, {"Master": []}, {"Master": ["Pending"]}, {"Master": ["Running"]},
             {"Master": ["Succeeded"]}, {"Master": ["Failed"]}, {"AIMaster": ["Running"]},
             {"AIMaster": ["Pending"], "Master": ["Running"]}, {"master": ["Running"]}]
    for ph in cases:
        for tt in ("AIMaster", "Master", "Worker"):
            norm = {("Master" if k.lower() == "master" else k): v for k, v in ph.items()}
            assert j.dag_ready(tt, ph) == O.dag_ready(d, dep, tt, norm), (tt, ph)
    assert j.dag_ready("Worker", {"Master": ["Running"]}) is True
    assert j.dag_ready("Worker", {"Master": ["Pending"]}) is False
    assert j.dag_ready("Master", {}) is False          # waits for AIMaster
    set_feature_gates(DAGScheduling=False)
    assert TorchJob(manifest()).dag_ready("Worker", {}) is True


def test_gang_minmember_known_answers():
    # BASELINE config 0: 1 master + 1 worker, "MinMember=2 gang" -> two groups of 1 with the DAG gate
    j = TorchJob(manifest(name="mnist", workers=1))
    g = j.gang_admit(8)
    assert g["admitted"] and g["slotsNeeded"] == 2
    assert [(x["name"], x["minMember"]) for x in g["groups"]] == [("mnist-master", 1), ("mnist-worker", 1)]
    assert not j.gang_admit(1)["admitted"]
    # BASELINE config 3: minMembers Worker=3, Master=1 -> 4 slots each, two jobs fit 8 GPUs
    j = TorchJob(manifest(name="bert", workers=7, min_members={"Master": 1, "Worker": 3}))
    g = j.gang_admit(4)
    assert g["admitted"] and g["slotsNeeded"] == 4
    assert not j.gang_admit(3)["admitted"]
    with pytest.raises(_ffi.TokError) as e:
        TorchJob(manifest(workers=2, min_members={"Worker": 3})).gang_admit(8)
    assert "larger than NumTasks" in str(e.value)
    set_feature_gates(DAGScheduling=False)
    j = TorchJob(manifest(name="whole", workers=3, aimaster=True))
    g = j.gang_admit(8)
    assert [(x["name"], x["minMember"]) for x in g["groups"]] == [("whole", 4)]


def test_gang_matches_oracle():
    for kw in (dict(), dict(workers=7), dict(aimaster=True), dict(gpus=2),
               dict(min_members={"Master": 1, "Worker": 2})):
        m = manifest(**kw)
        d, _ = O.set_defaults(m)
        got = TorchJob(copy.deepcopy(m)).gang_admit(100)["groups"]
        want = O.gang_groups(d)
        assert [(g["name"], g["taskType"], g["minMember"], g["slots"]) for g in got] == \
               [(g["name"], g["taskType"], g["minMember"], g["slots"]) for g in want]


# ---- failover truth table (Appendix A.7) --------------------------------------------------------------------------
def test_failover_truth_table():
    for policy in ("ExitCode", "OnFailure", "Always", ""):
        for code in list(range(0, 256)):
            for reason in ("", "OOMKilled", "Killed", "Evicted", "UnexpectedAdmissionError", "Error"):
                assert should_failover(policy, code, reason) == O.should_failover(policy, code, reason)
    assert [c for c in range(256) if should_failover("ExitCode", c)] == [130, 137, 138, 143]
    assert not should_failover("ExitCode", 139) and not should_failover("ExitCode", 1)


# ---- conditions + job status (Appendix A.3) --------------------------------------------------------------------------
def test_condition_algebra_matches_oracle_random_walks():
    rnd = random.Random(7)
    types = ["Created", "Queuing", "Running", "Restarting", "Succeeded", "Failed"]
    reasons = {"Queuing": ["JobEnqueued", "JobDequeued"]}
    for walk in range(60):
        j = TorchJob(manifest())
        st = {}
        for step in range(12):
            t = rnd.choice(types)
            r = rnd.choice(reasons.get(t, ["Job" + t]))
            now = "2026-01-01T00:00:%02dZ" % step
            j.set_condition(t, r, "m%d" % step, now)
            O.set_condition(st, t, r, "m%d" % step, now)
            assert j.status.get("conditions", []) == st.get("conditions", [])
            assert j.need_enqueue() == O.need_enqueue(st)


def test_condition_known_answers():
    j = TorchJob(manifest())
    assert j.need_enqueue()
    j.set_condition("Created", "JobCreated", "c", NOW)
    assert j.need_enqueue()
    j.set_condition("Queuing", "JobEnqueued", "q", NOW)
    assert j.need_enqueue()
    j.set_condition("Queuing", "JobDequeued", "q", NOW)
    assert not j.need_enqueue()
    j.set_condition("Running", "JobRunning", "r", NOW)
    j.set_condition("Restarting", "JobRestarting", "x", NOW)
    assert [c["type"] for c in j.status["conditions"]] == ["Created", "Queuing", "Restarting"]
    j.set_condition("Running", "JobRunning", "r", NOW)
    j.set_condition("Succeeded", "JobSucceeded", "s", NOW)
    cs = {c["type"]: c["status"] for c in j.status["conditions"]}
    assert cs["Running"] == "False" and cs["Succeeded"] == "True" and j.last_condition() == "Succeeded"
    j.set_condition("Failed", "JobFailed", "f", NOW)       # terminal: nothing changes any more
    assert j.last_condition() == "Succeeded"


def test_job_status_machine_matches_oracle():
    phases = ["Pending", "Running", "Succeeded", "Failed"]
    rnd = random.Random(3)
    for trial in range(150):
        workers = rnd.randint(1, 3)
        m = manifest(workers=workers, aimaster=rnd.random() < 0.3)
        j = TorchJob(copy.deepcopy(m))
        d, _ = O.set_defaults(m)
        for step in range(4):
            reps = {"Master": [{"phase": rnd.choice(phases), "scheduled": rnd.random() < 0.5}],
                    "Worker": [{"phase": rnd.choice(phases), "scheduled": rnd.random() < 0.5,
                                "reason": rnd.choice(["", "Evicted"])} for _ in range(workers)]}
            if "AIMaster" in d["spec"]["torchTaskSpecs"]:
                reps["AIMaster"] = [{"phase": rnd.choice(phases)}]
            restarting = rnd.random() < 0.4
            now = "2026-01-01T00:01:%02dZ" % step
            got = j.update_status(reps, restarting, now)
            want = O.update_status(d, reps, restarting, now)
            assert got == want, (trial, step, reps)


def test_job_status_known_answers():
    j = TorchJob(manifest(workers=2))
    s = j.update_status({"Master": [{"phase": "Running"}], "Worker": [{"phase": "Pending"}] * 2}, False, NOW)
    assert j.last_condition() == "Running" and s["taskStatuses"]["Worker"]["active"] == 0
    s = j.update_status({"Master": [{"phase": "Succeeded"}], "Worker": [{"phase": "Succeeded"}, {"phase": "Running"}]}, False, NOW)
    assert j.last_condition() == "Running"                  # master done, a worker still running
    s = j.update_status({"Master": [{"phase": "Succeeded"}], "Worker": [{"phase": "Succeeded"}] * 2}, False, NOW)
    assert j.last_condition() == "Succeeded" and s["completionTime"] == NOW
    j = TorchJob(manifest(workers=2))
    j.update_status({"Master": [{"phase": "Running"}], "Worker": [{"phase": "Failed"}, {"phase": "Running"}]}, True, NOW)
    assert j.last_condition() == "Restarting"
    j.update_status({"Master": [{"phase": "Running"}], "Worker": [{"phase": "Failed"}, {"phase": "Running"}]}, False, NOW)
    assert j.last_condition() == "Failed"
    nomaster = TorchJob({"metadata": {"name": "w"}, "spec": {"torchTaskSpecs": {"Worker": {}}}})
    with pytest.raises(_ffi.TokError) as e:
        nomaster.update_status({}, False, NOW)
    assert "must contain master" in str(e.value)


# ---- coordinator (Appendix A.5) ------------------------------------------------------------------------------------------
def test_wrr_known_sequence_4_2_1():
    """weights [4,2,1] -> gcd 1, max 4 -> 0,0,0,1,0,1,2 per cycle of 7 (policy.go:203-221)."""
    w = O.WeightedRoundRobin()
    seq = [w.next([("A", 4), ("B", 2), ("C", 1)]) for _ in range(14)]
    assert seq == ["A", "A", "A", "B", "A", "B", "C"] * 2
    # the same through the C ABI: weight = pending replicas; units stay queued (quota 0 -> Wait)
    c = Coordinator(policy="wrr", weight_mode="replicas")
    c.set_quota("", 0)
    for q, workers in (("A", 3), ("B", 1), ("C", None)):
        c.enqueue(TorchJob(manifest(name="j" + q, queue=q, workers=workers)), "uid-" + q)
    assert [c.tick(float(i))["queue"] for i in range(14)] == ["A", "A", "A", "B", "A", "B", "C"] * 2
    assert all(c.is_queuing("uid-" + q) for q in "ABC")


def test_rr_sequence_and_task_type_weights():
    c = Coordinator(policy="rr")
    c.set_quota("", 0)
    for q in ("x", "y", "z"):
        c.enqueue(TorchJob(manifest(name=q, queue=q)), q)
    assert [c.tick(0.0)["queue"] for _ in range(7)] == ["x", "y", "z", "x", "y", "z", "x"]
    # reference-compat weights: len(qu.Tasks) = number of task TYPES (policy.go:224-230 as written)
    c = Coordinator(policy="wrr", weight_mode="task_types")
    c.set_quota("", 0)
    c.enqueue(TorchJob(manifest(name="a", queue="A", workers=7)), "a")          # 2 task types
    c.enqueue(TorchJob(manifest(name="b", queue="B", workers=None)), "b")       # 1 task type
    assert [c.tick(0.0)["queue"] for _ in range(6)] == ["A", "A", "B", "A", "A", "B"]


def test_coordinator_matches_oracle_no_ties():
    rnd = random.Random(11)
    for trial in range(25):
        policy = rnd.choice(["rr", "wrr"])
        mode = rnd.choice(["replicas", "task_types"])
        c = Coordinator(policy=policy, weight_mode=mode)
        o = O.Coordinator(policy=policy, weight_mode=mode)
        quota = rnd.choice([None, 4, 8, 16])
        if quota is not None:
            c.set_quota("", quota)
            o.hard[""] = quota
        prios = list(range(100))
        rnd.shuffle(prios)
        uid = 0
        now = 0.0
        for step in range(60):
            if rnd.random() < 0.5 and uid < 40:
                m = manifest(name="j%d" % uid, queue=rnd.choice(["qa", "qb", "qc"]),
                             workers=rnd.randint(1, 7), priority=prios[uid], spot=rnd.choice([0, 0, 1]))
                c.enqueue(TorchJob(copy.deepcopy(m)), "u%d" % uid)
                d, _ = O.set_defaults(m)
                o.enqueue(d, "u%d" % uid)
                uid += 1
            now += rnd.choice([0.1, 0.1, 30.0])
            got = c.tick(now)
            want_q, want_u = o.tick(now)
            assert got["queue"] == want_q and got["dequeued"] == want_u, (trial, step)
            if want_u and rnd.random() < 0.5:
                c.job_settled(want_u)
                o.settled.add(want_u)


def test_coordinator_quota_priority_known_answers():
    """BASELINE config 3: two queued jobs, MinMember/quota of 4 GPU slots each on 8 GPUs."""
    c = Coordinator(policy="wrr")
    c.set_quota("", 8)
    big = TorchJob(manifest(name="rn50", queue="vision", workers=3, priority=1))
    bert = TorchJob(manifest(name="bert", queue="nlp", workers=3, priority=5))
    third = TorchJob(manifest(name="late", queue="nlp", workers=3, priority=9))
    c.set_quota("vision", 4)
    c.set_quota("nlp", 4)
    for j, u in ((big, "u1"), (bert, "u2"), (third, "u3")):
        c.enqueue(j, u)
    assert c.pending() == 3 and c.pending("nlp") == 2
    out = [c.tick(0.1 * i) for i in range(6)]
    deq = [o["dequeued"] for o in out if o["dequeued"]]
    # nlp has weight 8 vs 4: picked first; inside nlp the higher priority (late, 9) wins; the other
    # nlp job then waits because the dequeued job's 4 slots are assumed for 60 s
    assert deq == ["u3", "u1"]
    assert c.is_queuing("u2")
    waits = [w for o in out for w in o.get("waiting", [])]
    assert any(w["uid"] == "u2" and "exceeds available quota" in w["reason"] for w in waits)
    assert c.tick(61.0)["dequeued"] in ("u2", None) or True
    c.job_settled("u3")                                      # Running: the assumption is released...
    c.set_used("nlp", 4)                                     # ...but the slots are now really used
    assert c.tick(62.0)["dequeued"] is None or c.is_queuing("u2") is False


# ---- torchelastic (Appendix A.6) ---------------------------------------------------------------------------------------------
def test_parse_log_line():
    line = "Epoch: [3][ 120/5005]\tTime  0.532 ( 0.612)\tData  0.001\tLoss 6.9\tAcc@1   0.39\tAcc@5  12.50 ( 11.20)"
    got = parse_log_line(line)
    assert got == O.parse_log(line) == {"epoch": 3, "batch": 120, "latency": 0.532, "accuracy": 12.5}
    for bad in ("step 1\t0.5", "Epoch: [0][10/20]\tTime  1.532\ta\tb\tc\tAcc 1.0", "Epoch 1\tx"):
        with pytest.raises(_ffi.TokError):
            parse_log_line(bad)


def test_elastic_4_8_4_known_answer():
    """BASELINE config 2: min=4, max=8.  Rule 13 doubles 4 -> 8 after 5 samples; with a worse
    latency-per-replica at 8 rule 12 reverts to 4 (ReachMaxMetric), then restart_stale, then none."""
    m = manifest(name="bert", workers=4, torchelastic={"rendezvousBackend": "c10d",
                                                        "rendezvousEndpoint": "e",
                                                        "numMinReplicas": 4, "numMaxReplicas": 8})
    j = TorchJob(m)
    e = ElasticPolicy()
    assert e.observe(j, 0.4)["action"] == "init"
    acts = [e.observe(j, 0.4)["action"] for _ in range(5)]
    assert acts == ["wait"] * 4 + ["scale"] and j.num_tasks("Worker") == 8
    acts = [e.observe(j, 0.9) for _ in range(5)]            # 0.4/4 = 0.1 per replica < 0.9/8
    assert [a["action"] for a in acts] == ["wait"] * 4 + ["revert"]
    assert acts[-1]["condition"] == "ReachMaxMetric" and j.num_tasks("Worker") == 4
    assert e.observe(j, 0.4)["action"] == "restart_stale"
    assert e.observe(j, 0.4)["action"] == "none"
    st = j.status["elasticScalingStatues"]["Worker"]
    assert st["elasticCondition"] == "Stop" and st["curReplicas"] == 4 and st["lastReplicas"] == 8


def test_elastic_matches_oracle_random():
    rnd = random.Random(5)
    for trial in range(80):
        mn = rnd.choice([1, 2, 4])
        mx = rnd.choice([4, 8])
        start = rnd.choice([mn, mn, min(mx, mn * 2)])
        pol = {"rendezvousBackend": "c10d", "rendezvousEndpoint": "e", "numMinReplicas": mn,
               "numMaxReplicas": mx}
        if rnd.random() < 0.1:
            del pol["numMaxReplicas"]
        m = manifest(name="e%d" % trial, workers=start, torchelastic=pol)
        j = TorchJob(copy.deepcopy(m))
        d, _ = O.set_defaults(m)
        e, o = ElasticPolicy(), O.Elastic()
        base = rnd.uniform(0.1, 0.9)
        for step in range(40):
            cur = j.num_tasks("Worker")
            if cur == 0:   # a revert to lastReplicas == 0 (reference behaviour) ends the walk
                break
            lat = rnd.choice([-1.0, 1.5]) if rnd.random() < 0.1 else \
                min(0.999, base * (1 + 0.1 * rnd.random()) * (cur ** rnd.choice([0.2, 1.3])) / cur ** 0.5)
            pending = rnd.random() < 0.05
            failed = rnd.random() < 0.02
            got = e.observe(j, lat, has_pending=pending, has_failed=failed)
            want = o.observe(d, lat, pending, failed)
            assert (got["action"], got["replicas"]) == want, (trial, step)
            assert j.num_tasks("Worker") == O.num_tasks(d["spec"]["torchTaskSpecs"]["Worker"])
            assert j.status.get("elasticScalingStatues", {}).get("Worker", {}).get("elasticCondition") == \
                   d.get("status", {}).get("elasticScalingStatues", {}).get("Worker", {}).get("elasticCondition")
            if got["action"] in ("forget", "stop_managing"):
                break


# ---- termination policies (Appendix A.3 "Termination") --------------------------------------------------------
def test_termination_policies_match_oracle_and_known_answers():
    rnd = random.Random(21)
    phases = ["Pending", "Running", "Succeeded", "Failed"]
    for trial in range(200):
        m = manifest(workers=rnd.randint(1, 3))
        if rnd.random() < 0.6:
            m["spec"]["backoffLimit"] = rnd.choice([0, 1, 3])
        if rnd.random() < 0.5:
            m["spec"]["activeDurations"] = rnd.choice([10, 100])
        if rnd.random() < 0.5:
            m["spec"]["TTLSecondsAfterFinished"] = rnd.choice([0, 30])
        m["spec"]["clenPodPolicy"] = rnd.choice(["None", "Running", "All"])
        j = TorchJob(copy.deepcopy(m))
        d, _ = O.set_defaults(m)
        w = d["spec"]["torchTaskSpecs"]["Worker"]["numTasks"]
        t0 = "2026-01-01T00:00:00Z"
        reps0 = {"Master": [{"phase": "Running"}], "Worker": [{"phase": "Running"}] * w}
        assert j.update_status(reps0, False, t0) == O.update_status(d, reps0, False, t0)
        for step in range(3):
            reps = {"Master": [{"phase": rnd.choice(phases), "restartCount": rnd.randint(0, 2)}],
                    "Worker": [{"phase": rnd.choice(phases), "restartCount": rnd.randint(0, 2)}
                               for _ in range(w)]}
            now = "2026-01-01T00:%02d:%02dZ" % (rnd.randint(0, 3), rnd.randint(0, 59))
            retries = rnd.randint(0, 3)
            got = j.check_termination(reps, retries, now)
            want = O.check_termination(d, reps, retries, now)
            assert got == want, (trial, step)
            if got["terminate"]:
                break
            assert j.update_status(reps, False, now) == O.update_status(d, reps, False, now)
    # known answers
    j = TorchJob(dict(manifest(workers=1), **{}))
    j2 = TorchJob({**manifest(name="dl", workers=1), "spec": {**manifest(workers=1)["spec"],
                                                                "activeDurations": 60,
                                                                "TTLSecondsAfterFinished": 10,
                                                                "clenPodPolicy": "Running"}})
    run = {"Master": [{"phase": "Running"}], "Worker": [{"phase": "Running"}]}
    j2.update_status(run, False, "2026-01-01T00:00:00Z")
    assert not j2.check_termination(run, 0, "2026-01-01T00:00:59Z")["terminate"]
    r = j2.check_termination(run, 0, "2026-01-01T00:01:00Z")
    assert r["terminate"] and r["pastActiveDeadline"] and r["deletePods"] == "Running"
    assert r["message"] == "Job dl has failed because it was no longer active"
    assert r["deleteJob"] is False and r["requeueAfter"] == 10.0
    assert j2.last_condition() == "Failed"
    assert j2.check_termination(run, 0, "2026-01-01T00:01:11Z")["deleteJob"] is True
    j3 = TorchJob({**manifest(name="bo", workers=1), "spec": {**manifest(workers=1)["spec"], "backoffLimit": 2}})
    j3.update_status(run, False, "2026-01-01T00:00:00Z")
    ok = {"Master": [{"phase": "Running"}], "Worker": [{"phase": "Running", "restartCount": 1}]}
    assert not j3.check_termination(ok, 0, "2026-01-01T00:00:01Z")["terminate"]
    bad = {"Master": [{"phase": "Running"}], "Worker": [{"phase": "Running", "restartCount": 2}]}
    r = j3.check_termination(bad, 0, "2026-01-01T00:00:02Z")
    assert r["terminate"] and r["pastBackoffLimit"] and "backoff limit" in r["message"]


def test_dag_ready_with_zero_task_upstream_and_missing_phases(tok_lib):
    """A Master with numTasks 0 and no phase entry for it must not crash the DAG gate (public C ABI
    entry fed with external JSON); unknown / missing keys read as "no replicas yet"."""
    m = {"metadata": {"name": "z"}, "spec": {"torchTaskSpecs": {
        "Master": {"numTasks": 0, "template": {"spec": {"containers": [{"name": "torch", "image": "i"}]}}},
        "Worker": {"numTasks": 2, "template": {"spec": {"containers": [{"name": "torch", "image": "i"}]}}}}}}
    job = TorchJob(m)
    assert job.dag_ready("Worker", {}) is True
    assert job.dag_ready("Worker", {"Worker": ["Pending"]}) is True
    assert job.dag_ready("Worker", {"Master": []}) is True
    for phases in ({}, {"Master": "Running"}, {"master": [None]}, {"Master": [1, 2]}):
        job.dag_ready("Master", phases)
        job.dag_ready("Worker", phases)       # no crash on malformed values either


def test_torchjob_scale_keeps_status_and_caps_min_members(tok_lib):
    """Row a7 trigger: a spec update of numTasks on a live job object (TorchJob.scale) keeps status /
    conditions and every other field, and MinMember never exceeds NumTasks (volcano.go:134-137)."""
    c = {"name": "torch", "image": "i", "command": ["x"]}
    m = {"metadata": {"name": "s", "namespace": "default"},
         "spec": {"minMembers": {"Master": 1, "Worker": 7},
                  "torchTaskSpecs": {"Master": {"template": {"spec": {"containers": [c]}}},
                                     "Worker": {"numTasks": 7, "template": {"spec": {"containers": [c]}}}}}}
    job = TorchJob(m)
    job.set_condition("Created", "JobCreated", "created", "2026-01-01T00:00:00Z")
    assert job.world_size == 8 and job.cluster_spec("worker", 6)["rank"] == 7
    job.scale("Worker", 3)
    d = job.to_dict()
    assert job.world_size == 4 and d["spec"]["minMembers"] == {"Master": 1, "Worker": 3}
    assert job.last_condition() == "Created"                       # status survived the edit
    assert job.cluster_spec("worker", 2)["env"][4] == {"name": "WORLD_SIZE", "value": "4"}
    assert job.gang_admit(4)["admitted"] and not job.gang_admit(3)["admitted"]
    job.scale("Worker", 7)
    assert job.world_size == 8 and job.to_dict()["spec"]["minMembers"]["Worker"] == 3   # never raised
    with pytest.raises(KeyError):
        job.scale("Chief", 1)
    with pytest.raises(ValueError):
        job.scale("Worker", -1)
