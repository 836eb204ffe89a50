"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by torch-on-k8s_b200/).

Pure-Python restatement of the reference's control-plane functions ON the hot path's configuration
side.  The reference is Go and cannot be compiled or run here (no Go
toolchain) and ships no tests, so this restatement is "parity unpinned" by the
reference: it is authored from the sources cited per function and from the behaviour tables in
SURVEY.md Appendix A, and the C++ implementation behind the C ABI (csrc/ctl_*.cpp) is tested
against it plus hand-derived known answers (tests/test_controlplane.py).

Where the reference has a latent defect, the INTENDED behaviour of SURVEY.md §2.3 is restated (and
marked), exactly as the product does.
"""
from __future__ import annotations

import copy
import math
import re
from typing import Dict, List, Optional, Tuple

GATES_DEFAULT = {"GangScheduling": True, "DAGScheduling": True, "JobCoordinator": True,
                 "TorchLocalMasterAddr": True, "HostNetWithHeadlessSvc": False}
# pkg/features/features.go:54-63


def num_tasks(ts: dict) -> int:
    n = ts.get("numTasks")
    return 1 if n is None else int(n)


def gen_general_name(job: str, task_type: str, index) -> str:
    """pkg/utils/utils.go:75-77"""
    return ("%s-%s-%s" % (job, task_type, index)).replace("/", "-")


# ---- apis/train/v1alpha1/torchjob_defaults.go:29-74 ----------------------------------------------
def set_defaults(job: dict, gates: Optional[dict] = None) -> Tuple[dict, dict]:
    """Returns (defaulted job, DependsOn map) — DependsOn is `json:"-"` (torchjob_types.go:103)."""
    g = dict(GATES_DEFAULT, **(gates or {}))
    job = copy.deepcopy(job)
    spec = job.setdefault("spec", {})
    if spec.get("clenPodPolicy") is None:                     # :31-34
        spec["clenPodPolicy"] = "None"
    specs = spec["torchTaskSpecs"]
    for canon in ("Master", "Worker"):                        # :77-92
        for k in list(specs):
            if k.lower() == canon.lower() and k != canon and canon not in specs:
                specs[canon] = specs.pop(k)
                break
    depends: Dict[str, List[Tuple[str, str]]] = {}
    if g["DAGScheduling"]:                                    # :95-124
        if "AIMaster" in specs and "Master" in specs:
            depends["Master"] = [("AIMaster", "Running")]
        if "Worker" in specs and "Master" in specs:
            depends["Worker"] = [("Master", "Running")]
    for tt, ts in specs.items():
        if tt == "Worker":                                    # :140-147
            if ts.get("numTasks") is None:
                ts["numTasks"] = 1
            if not ts.get("restartPolicy"):
                ts["restartPolicy"] = "OnFailure"
        if tt == "Master":                                    # :128-136, 150-178
            if ts.get("numTasks") is None:
                ts["numTasks"] = 1
            if not ts.get("restartPolicy"):
                ts["restartPolicy"] = "ExitCode"
            for c in ts.get("template", {}).get("spec", {}).get("containers", []) or []:
                if c.get("name") == "torch":
                    ports = c.setdefault("ports", [])
                    if not any(p.get("name") == "torchjob-port" for p in ports):
                        ports.append({"name": "torchjob-port", "containerPort": 23456})
                    break
        for c in ts.get("template", {}).get("spec", {}).get("containers", []) or []:   # :181-188
            if not c.get("terminationMessagePolicy"):
                c["terminationMessagePolicy"] = "FallbackToLogsOnError"
    if not job.get("apiVersion"):                             # :60-65
        job["apiVersion"] = "train.distributed.io/v1alpha1"
    if not job.get("kind"):
        job["kind"] = "TorchJob"
    # :69-73, 192-197 — INTENDED: the reference iterates the nil map it guards (no-op)
    if g["DAGScheduling"] and g["GangScheduling"] and spec.get("minMembers") is None:
        spec["minMembers"] = {tt: num_tasks(ts) for tt, ts in specs.items()}
    return job, depends


# ---- controllers/train/torchjob_controller.go:314-449 ----------------------------------------------
def master_port(specs: dict) -> int:
    """getPortFromJob (:508-521)"""
    for c in specs["Master"].get("template", {}).get("spec", {}).get("containers", []) or []:
        if c.get("name") == "torch":
            for p in c.get("ports", []) or []:
                if p.get("name") == "torchjob-port":
                    return int(p["containerPort"])
    raise ValueError("failed to found the port")


def cluster_spec(job: dict, task_type: str, index: int, gates: Optional[dict] = None) -> dict:
    g = dict(GATES_DEFAULT, **(gates or {}))
    specs = job["spec"]["torchTaskSpecs"]
    name = job["metadata"]["name"]
    tt = task_type.lower()
    key = next(k for k in specs if k.lower() == tt)
    port = master_port(specs)
    master_role = tt == "master"
    master_addr = gen_general_name(name, "master", 0)        # :338
    rank = index
    if master_role:
        if rank != 0:                                        # :339-342
            raise ValueError("invalid config: There should be only a single master with index=0")
        if g["TorchLocalMasterAddr"]:
            master_addr = "localhost"
    else:
        rank += 1                                            # :346-348
    world = sum(num_tasks(ts) for k, ts in specs.items() if k != "AIMaster")   # :350
    elastic = (job["metadata"].get("annotations") or {}).get(
        "distributed.io/enable-elastic-training") == "true"
    env = [("MASTER_PORT", str(port)), ("MASTER_ADDR", master_addr), ("RANK", str(rank)),
           ("PYTHONUNBUFFERED", "0")]                         # :398-413
    annotations, finalizers, init = {}, [], []
    task_rp = specs[key].get("restartPolicy", "")
    pod_rp = "Never" if task_rp == "ExitCode" else task_rp   # controllers/common/pod.go:556-561
    labels = {"group-name": "train.distributed.io", "job-name": name.replace("/", "-"),
              "task-type": tt, "task-index": str(index)}     # pod.go:519-524
    if master_role:
        labels["task-role"] = "master"
    if elastic and tt != "aimaster":                          # :419-439
        annotations["distributed.io/world-size"] = str(world)
        env.append(("WORLD_SIZE", str(world)))
        pod_rp = "OnFailure"
        labels["distributed.io/job-generation"] = str(job["metadata"].get("generation", 0))
        finalizers.append("distributed.io/preempt-protector")
        if not master_role:                                   # :351-361
            init = ["warmup", "master-waiter"]
    else:
        env.append(("WORLD_SIZE", str(world)))
    args = []
    pol = job["spec"].get("torchElasticPolicy")
    if job["spec"].get("enableTorchElastic") and pol:         # :415 guard; :364-392
        desired = num_tasks(specs["Worker"]) if "Worker" in specs else 1   # INTENDED (§2.3)
        vmin = pol.get("numMinReplicas", None)
        vmax = pol.get("numMaxReplicas", None)
        nproc = pol.get("numWorkersPerNodePolicy", None)
        args = ["--rdzv_backend=" + pol.get("rendezvousBackend", ""),
                "--rdzv_endpoint=" + pol.get("rendezvousEndpoint", ""),
                "--rdzv_id=" + name,
                "--nproc_per_node=%d" % (1 if nproc is None else nproc),
                "--nnodes=%d:%d" % (desired if vmin is None else vmin,
                                    desired if vmax is None else vmax)]
    if g["GangScheduling"] and tt != "aimaster":              # volcano.go:238-287
        annotations["scheduling.k8s.io/group-name"] = (name + "-" + tt) if g["DAGScheduling"] else name
    return dict(name=gen_general_name(name, tt, index), rank=rank, worldSize=world, env=env,
                args=args, labels=labels, annotations=annotations, restartPolicy=pod_rp,
                initContainers=init, finalizers=finalizers)


# ---- controllers/common/dag.go:30-116 --------------------------------------------------------------
PHASE_CODE = {"Pending": 0, "Running": 1, "Succeeded": 2, "Failed": 2}


def dag_ready(job: dict, depends: dict, task_type: str, phases: Dict[str, List[str]]) -> bool:
    specs = job["spec"]["torchTaskSpecs"]
    for upstream, on_phase in depends.get(task_type, []):
        if upstream not in specs:
            continue
        have = phases.get(upstream, [])
        if len(have) < num_tasks(specs[upstream]):
            return False
        if any(PHASE_CODE.get(p, 0) - PHASE_CODE[on_phase] < 0 for p in have):
            return False
    return True


# ---- pkg/gangscheduler/volcano/volcano.go:109-230 ----------------------------------------------------
def replica_slots(tt: str, ts: dict) -> int:
    s = 0
    for c in ts.get("template", {}).get("spec", {}).get("containers", []) or []:
        r = (c.get("resources") or {})
        v = (r.get("requests") or {}).get("nvidia.com/gpu", (r.get("limits") or {}).get("nvidia.com/gpu"))
        if v is not None:
            s += int(v)
    if s == 0 and tt != "AIMaster":
        s = 1
    return s


def gang_groups(job: dict, gates: Optional[dict] = None) -> List[dict]:
    g = dict(GATES_DEFAULT, **(gates or {}))
    specs = job["spec"]["torchTaskSpecs"]
    name = job["metadata"]["name"]
    if not g["GangScheduling"]:
        return []
    if g["DAGScheduling"]:                                    # generatePodGroupsByRole
        out = []
        mm = job["spec"].get("minMembers") or {}
        for tt, ts in specs.items():
            if tt == "AIMaster":
                continue
            n = num_tasks(ts)
            m = mm.get(tt)
            if m is not None and m > n:
                raise ValueError("the mimMember provided for task type %s is larger than NumTasks" % tt)
            m = n if m is None else m
            out.append(dict(name="%s-%s" % (name, tt.lower()), taskType=tt, minMember=m,
                            slots=m * replica_slots(tt, ts)))
        return out
    total = sum(num_tasks(ts) for tt, ts in specs.items() if tt != "AIMaster")
    slots = sum(num_tasks(ts) * replica_slots(tt, ts) for tt, ts in specs.items() if tt != "AIMaster")
    ma = ((job["spec"].get("schedulingPolicy") or {}).get("minAvailable"))
    m = total
    if ma is not None and ma > 0:
        m = ma
        slots = slots * m // total if total else slots
    return [dict(name=name, taskType="", minMember=m, slots=slots)]


# ---- controllers/common/failover.go:52-113 ------------------------------------------------------------
def should_failover(restart_policy: str, exit_code: int, reason: str = "") -> bool:
    if restart_policy != "ExitCode":
        return False
    retryable = exit_code in (130, 137, 143, 138)
    return retryable or reason in ("OOMKilled", "Killed", "Evicted", "UnexpectedAdmissionError")


# ---- pkg/utils/utils.go:186-243 -------------------------------------------------------------------------
def has_condition(status: dict, ctype: str) -> bool:
    return any(c["type"] == ctype and c["status"] == "True" for c in status.get("conditions", []))


def set_condition(status: dict, ctype: str, reason: str, message: str, now: str) -> None:
    if has_condition(status, "Failed") or has_condition(status, "Succeeded"):
        return
    conds = status.setdefault("conditions", [])
    cur = next((c for c in conds if c["type"] == ctype), None)
    transition = now
    if cur is not None and cur["status"] == "True":
        if cur["reason"] == reason:
            return
        transition = cur["lastTransitionTime"]
    kept = []
    for c in conds:
        if ctype == "Restarting" and c["type"] == "Running":
            continue
        if ctype == "Running" and c["type"] == "Restarting":
            continue
        if c["type"] == ctype:
            continue
        if ctype in ("Failed", "Succeeded") and c["type"] == "Running":
            c = dict(c, status="False")
        kept.append(c)
    kept.append(dict(type=ctype, status="True", lastUpdateTime=now, lastTransitionTime=transition,
                     reason=reason, message=message))
    status["conditions"] = kept


def need_enqueue(status: dict) -> bool:
    """utils.go:138-149"""
    conds = status.get("conditions", [])
    if not conds:
        return True
    last = conds[-1]
    return last["type"] == "Created" or (last["type"] == "Queuing" and last["reason"] == "JobEnqueued")


# ---- controllers/common/pod.go:690-714 + controllers/train/job.go:99-207 ---------------------------------
def update_status(job: dict, replicas: Dict[str, List[dict]], restarting: bool, now: str) -> dict:
    specs = job["spec"]["torchTaskSpecs"]
    name = job["metadata"]["name"]
    status = job.setdefault("status", {})
    ts_all = {}
    for tt in specs:
        a = s = f = ev = 0
        for r in replicas.get(tt, []):
            ph = r.get("phase")
            if ph == "Pending":
                if r.get("scheduled", False) and r.get("initPassed", True):
                    a += 1
            elif ph == "Running":
                a += 1
            elif ph == "Succeeded":
                s += 1
            elif ph == "Failed":
                f += 1
                ev += r.get("reason") == "Evicted"
        ts_all[tt] = dict(active=a, succeed=s, failed=f, **({"evicted": ev} if ev else {}))
    status["taskStatuses"] = ts_all
    if not status.get("startTime"):
        status["startTime"] = now
    if "Master" not in specs and "AIMaster" not in specs:
        raise ValueError("invalid config: Job must contain master replica spec")
    worker = specs.get("Worker")
    all_workers = worker is not None and num_tasks(worker) == ts_all["Worker"]["succeed"]
    order = [k for k in ("AIMaster", "Master", "Worker") if k in specs] + \
            [k for k in specs if k not in ("AIMaster", "Master", "Worker")]
    for tt in order:
        n = num_tasks(specs[tt])
        t = ts_all[tt]
        expected = n - t["succeed"]
        if tt in ("Master", "AIMaster"):
            if t["active"] > 0:
                set_condition(status, "Running", "JobRunning", "TorchJob %s is running." % name, now)
            succeed = n > 0 and expected == 0
            if tt != "AIMaster" and worker is not None:
                succeed = succeed and all_workers
            if succeed:
                status.setdefault("completionTime", now)
                set_condition(status, "Succeeded", "JobSucceeded",
                              "TorchJob %s is successfully completed." % name, now)
        if t["failed"] > 0:
            if restarting and tt != "AIMaster":
                set_condition(status, "Restarting", "JobRestarting",
                              "TorchJob %s is restarting because %d %s task(s) failed." %
                              (name, t["failed"], tt), now)
            else:
                status.setdefault("completionTime", now)
                set_condition(status, "Failed", "JobFailed",
                              "TorchJob %s is failed because %d %s task(s) failed." %
                              (name, t["failed"], tt), now)
    return status


# ---- pkg/coordinator/core/policy.go -----------------------------------------------------------------------
class RoundRobin:
    """policy.go:31-76"""

    def __init__(self):
        self.names: List[str] = []
        self.last = -1

    def next(self, queue_names: List[str]) -> Optional[str]:
        if not self.names or len(self.names) != len(queue_names):
            self.names += [q for q in queue_names if q not in self.names]
        if not self.names:
            return None
        name = self.names[(self.last + 1) % len(self.names)]
        self.last += 1
        return name


class WeightedRoundRobin:
    """policy.go:80-230 with its maps initialised (§2.3); state (cur, cw) survives weight updates."""

    def __init__(self):
        self.names: List[str] = []
        self.weights: List[int] = []
        self.cur = -1
        self.cw = 0

    def next(self, queues: List[Tuple[str, int]]) -> Optional[str]:
        for name, w in queues:
            if name not in self.names:
                self.names.append(name)
                self.weights.append(w)
            else:
                self.weights[self.names.index(name)] = w
        if not self.names:
            return None
        g = 0
        for w in self.weights:
            g = math.gcd(g, w)
        mx = max(self.weights)
        if mx <= 0:      # every queue drained (the reference would spin forever: gcd == 0)
            return None
        n = len(self.weights)
        while True:                                            # nextQueueIndex, :203-221
            self.cur = (self.cur + 1) % n
            if self.cur == 0:
                self.cw -= g
                if self.cw <= 0:
                    self.cw = mx
                    if self.cw == 0:
                        return None
            if self.weights[self.cur] >= self.cw:
                return self.names[self.cur]


class Coordinator:
    """core/coordinator.go:164-476 + plugins/quota.go + plugins/priority.go over GPU slots.
    Ties are broken by the caller-supplied `tie(n) -> int in [0,n)` so tests can inject the same
    deterministic stream the C++ side uses (or avoid ties)."""

    def __init__(self, policy="wrr", weight_mode="replicas", tie=None):
        self.queues: Dict[str, List[dict]] = {}
        self.sel = WeightedRoundRobin() if policy == "wrr" else RoundRobin()
        self.policy = policy
        self.weight_mode = weight_mode
        self.hard: Dict[str, int] = {}
        self.used: Dict[str, int] = {}
        self.assumed: Dict[str, Dict[str, Tuple[int, float]]] = {}
        self.settled = set()
        self.tie = tie or (lambda n: 0)

    @staticmethod
    def unit(job: dict, uid: str) -> dict:
        sp = job["spec"].get("schedulingPolicy") or {}
        ns = job["metadata"].get("namespace") or "default"
        specs = job["spec"]["torchTaskSpecs"]
        slots = 0
        for tt, ts in specs.items():
            n = num_tasks(ts)
            spot = (ts.get("spotTaskSpec") or {}).get("numSpotTasks", 0)
            if spot > 0:
                n = max(0, n - spot)
            slots += n * replica_slots(tt, ts)
        return dict(uid=uid, tenant=sp.get("queue") or ns, priority=sp.get("priority"), slots=slots,
                    task_types=len(specs), replicas=sum(num_tasks(ts) for ts in specs.values()))

    def enqueue(self, job: dict, uid: str):
        u = self.unit(job, uid)
        q = self.queues.setdefault(u["tenant"], [])
        self.settled.discard(uid)
        for i, e in enumerate(q):
            if e["uid"] == uid:
                q[i] = u
                return
        q.append(u)

    def weight(self, q: List[dict]) -> int:
        return sum(u["task_types"] if self.weight_mode == "task_types" else u["replicas"] for u in q)

    def quota_key(self, u: dict) -> str:
        """usage / assumptions are kept under the quota object the hard limit came from: the tenant's
        own, else the default "" (the whole box) shared by every tenant without one"""
        return u["tenant"] if u["tenant"] in self.hard else ""

    def quota_ok(self, u: dict, now: float) -> bool:
        qk = self.quota_key(u)
        hard = self.hard.get(qk)
        if hard is None:
            return True
        used = self.used.get(qk, 0)
        if used > hard:
            return False
        a = self.assumed.setdefault(qk, {})
        for k in [k for k, (s, ts) in a.items() if now - ts > 60.0 or k in self.settled]:
            del a[k]
        avail = max(0, hard - used - sum(s for s, _ in a.values()))
        return avail >= u["slots"]

    def tick(self, now: float) -> Tuple[Optional[str], Optional[str]]:
        names = list(self.queues)
        if self.policy == "wrr":
            tenant = self.sel.next([(n, self.weight(self.queues[n])) for n in names])
        else:
            tenant = self.sel.next(names)
        if tenant is None:
            return None, None
        cands = [(u, u["priority"] or 0) for u in self.queues[tenant] if self.quota_ok(u, now)]
        if not cands:
            return tenant, None
        best, sel, ties = cands[0][1], 0, 1
        for i in range(1, len(cands)):
            if cands[i][1] > best:
                best, sel, ties = cands[i][1], i, 1
            elif cands[i][1] == best:
                ties += 1
                if self.tie(ties) == 0:
                    sel = i
        u = cands[sel][0]
        self.assumed.setdefault(self.quota_key(u), {})[u["uid"]] = (u["slots"], now)
        self.queues[tenant] = [e for e in self.queues[tenant] if e["uid"] != u["uid"]]
        return tenant, u["uid"]


# ---- controllers/train/torchelastic ---------------------------------------------------------------------------
def parse_log(line: str) -> dict:
    """observation.go:54-76"""
    f = line.rstrip("\r\n").split("\t")
    if "Epoch" not in f[0]:
        raise ValueError("current line of log is not a torchelastic training log")
    epoch = int(re.search(r"[0-9]{1,2}", f[0]).group(0))
    batch = int(re.search(r"[0-9]{2,4}", f[0]).group(0))
    lat = float(re.search(r"[0-9]{1,2}.[0-9]{3}", f[1]).group(0))
    acc = float(re.search(r"[0-9]{1,2}.[0-9]{1,2}", f[5]).group(0))
    if lat > 1:
        raise ValueError("epoch training time > 1, drop it")
    return dict(epoch=epoch, batch=batch, latency=lat, accuracy=acc)


class Elastic:
    """elastic_scale.go:42-246 (decision table of SURVEY.md Appendix A.6), doubling clamped to max."""

    def __init__(self, metric_count=5):
        self.k = metric_count
        self.metrics: Dict[int, List[float]] = {}

    def observe(self, job: dict, latency: float, pending=False, failed=False) -> Tuple[str, int]:
        w = job["spec"]["torchTaskSpecs"]["Worker"]
        pol = job["spec"].get("torchElasticPolicy") or {}
        mn, mx = pol.get("numMinReplicas"), pol.get("numMaxReplicas")
        cur = num_tasks(w)
        if mn is None or mx is None:
            return "forget", cur
        st = job.setdefault("status", {})
        es = st.get("elasticScalingStatues")
        if not es:
            st["elasticScalingStatues"] = {"Worker": {"elasticCondition": "Start", "continue": True,
                                                      "curReplicas": cur}}
            return "init", cur
        ws = es["Worker"]
        if st.get("completionTime") or job["metadata"].get("deletionTimestamp"):
            self.metrics = {}
            return "forget", cur
        last = ws.get("lastReplicas", 0)
        if pending and cur > mn:
            w["numTasks"] = last
            ws.update(elasticCondition="Stop", **{"continue": False}, lastReplicas=ws.get("curReplicas", cur),
                      curReplicas=last)
            return "revert", last
        if (pending and cur == mn) or failed:
            return "stop_managing", cur
        if not pending and not ws.get("continue", False):
            if ws.get("elasticCondition") == "ReachMaxMetric":
                ws["elasticCondition"] = "Stop"
                return "restart_stale", cur
            if ws.get("elasticCondition") in ("Stop", "ReachMaxReplicas"):
                return "none", cur
        if latency < 0 or latency > 1.0:
            return "skip", cur
        self.metrics.setdefault(cur, []).append(latency)
        if len(self.metrics[cur]) < self.k:
            return "wait", cur

        def up():
            nxt = min(cur * 2, mx)
            w["numTasks"] = nxt
            ws.update(elasticCondition="Continue", **{"continue": True}, lastReplicas=cur, curReplicas=nxt)
            self.metrics.setdefault(nxt, [])
            return "scale", nxt
        if mn < cur <= mx:
            base = self.metrics.get(last, [])
            better = len(base) < self.k or base[self.k - 1] / last > self.metrics[cur][self.k - 1] / cur
            if better:
                if cur == mx:
                    ws.update(elasticCondition="ReachMaxReplicas", **{"continue": False})
                    self.metrics[cur] = []
                    return "none", cur
                return up()
            w["numTasks"] = last
            ws.update(elasticCondition="ReachMaxMetric", **{"continue": False}, curReplicas=last,
                      lastReplicas=cur)
            self.metrics[last] = []
            self.metrics[cur] = []
            return "revert", last
        if cur == mn and cur < mx:
            return up()
        if cur == mx:
            ws.update(elasticCondition="ReachMaxReplicas", **{"continue": False})
            self.metrics[cur] = []
            return "none", cur
        return "none", cur


# ---- controllers/common/job.go:100-200, 385-460, 511-539 (termination policies) ------------------------------
def _ts(t: str) -> float:
    import calendar
    import time
    return float(calendar.timegm(time.strptime(t[:19], "%Y-%m-%dT%H:%M:%S")))


def check_termination(job: dict, replicas: Dict[str, List[dict]], prev_retries: int, now: str) -> dict:
    specs = job["spec"]["torchTaskSpecs"]
    spec = job["spec"]
    status = job.setdefault("status", {})
    name = job["metadata"]["name"]
    expected = sum(num_tasks(ts) for ts in specs.values())                 # GetTotalTasks
    active = failed = restarts = 0
    for tt, ts in specs.items():
        for r in replicas.get(tt, []):
            ph = r.get("phase")
            active += ph in ("Pending", "Running")                          # IsPodActive
            failed += ph == "Failed"
            if ph == "Running" and ts.get("restartPolicy") in ("OnFailure", "Always"):
                restarts += r.get("restartCount", 0)                        # pastBackoffLimit
    prev_failed = sum(t.get("failed", 0) for t in (status.get("taskStatuses") or {}).values())
    exceeds = past = deadline = False
    limit = spec.get("backoffLimit")
    if limit is not None:
        exceeds = failed > prev_failed and active != expected and prev_retries + 1 > limit
        past = restarts > 0 if limit == 0 else restarts >= limit
    msg = None
    if exceeds or past:
        msg = "Job %s has failed because it has reached the specified backoff limit" % name
    elif spec.get("activeDurations") is not None and status.get("startTime") and \
            _ts(now) - _ts(status["startTime"]) >= spec["activeDurations"]:
        deadline = True
        msg = "Job %s has failed because it was no longer active" % name
        status.setdefault("completionTime", now)   # INTENDED: once (the reference re-stamps it every pass)
    terminate = has_condition(status, "Succeeded") or has_condition(status, "Failed") or msg is not None
    res = dict(terminate=terminate, exceedsBackoffLimit=exceeds, pastBackoffLimit=past,
               pastActiveDeadline=deadline)
    if terminate:
        pol = spec.get("clenPodPolicy") or "None"
        res["deletePods"] = "None" if pol == "None" else ("Running" if pol == "Running" else "All")
        if msg is not None:
            status.setdefault("completionTime", now)
            set_condition(status, "Failed", "JobFailed", msg, now)
            res["message"] = msg
        if has_condition(status, "Succeeded"):
            for t in (status.get("taskStatuses") or {}).values():
                t["succeed"] = t.get("succeed", 0) + t.get("active", 0)
                t["active"] = 0
        ttl = spec.get("TTLSecondsAfterFinished")
        if ttl is not None and status.get("completionTime"):
            delete_at = _ts(status["completionTime"]) + ttl
            res["deleteJob"] = _ts(now) > delete_at
            res["requeueAfter"] = 0.0 if _ts(now) > delete_at else delete_at - _ts(now)
    res["status"] = status
    return res
