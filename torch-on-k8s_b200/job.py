"""TorchJob surface over the C ABI (csrc/ctl_job.cpp): the host-side mirror of the reference's
operator interface for the hot path's configuration half — same names, argument meaning and error
behaviour as apis/train/v1alpha1 + controllers/train (see include/tok8s.h for file:line cites)."""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, List, Optional, Union

from ._ffi import call_json, check, lib

TASK_ORDER = ["AIMaster", "Master", "Worker"]  # GetTaskReconcilerOrders, torchjob_controller.go:464-471


def load_manifest(text_or_path: str) -> dict:
    """JSON or YAML TorchJob manifest -> dict (YAML is converted on the host; the C ABI takes JSON)."""
    text = text_or_path
    if "\n" not in text_or_path and not text_or_path.lstrip().startswith("{"):
        with open(text_or_path) as f:
            text = f.read()
    try:
        return json.loads(text)
    except ValueError:
        import yaml
        return yaml.safe_load(text)


class TorchJob:
    """A parsed TorchJob (train.distributed.io/v1alpha1)."""

    def __init__(self, manifest: Union[str, dict], *, apply_defaults: bool = True):
        if not isinstance(manifest, str):
            manifest = json.dumps(manifest)
        self._h = C.c_void_p()
        check(lib().tok_job_parse(manifest.encode(), C.byref(self._h)))
        if apply_defaults:
            self.set_defaults()

    # SetDefaults_TorchJob
    def set_defaults(self) -> "TorchJob":
        check(lib().tok_job_default(self._h))
        return self

    def to_dict(self) -> dict:
        return call_json(lib().tok_job_to_json, self._h)

    @property
    def name(self) -> str:
        return self.to_dict()["metadata"]["name"]

    @property
    def task_specs(self) -> Dict[str, dict]:
        return self.to_dict()["spec"]["torchTaskSpecs"]

    def num_tasks(self, task_type: str) -> int:
        ts = self.task_specs.get(task_type)
        return 0 if ts is None else int(ts.get("numTasks", 1))

    @property
    def world_size(self) -> int:
        return sum(int(ts.get("numTasks", 1)) for tt, ts in self.task_specs.items() if tt != "AIMaster")

    def scale(self, task_type: str, num_tasks: int) -> None:
        """Edit spec.torchTaskSpecs[task_type].numTasks in place (`kubectl scale` / a spec update —
        the trigger of the reference's annotation-driven rescale, controllers/train/elastic_scale.go:
        210-397); status and every other field are preserved."""
        d = self.to_dict()
        specs = d["spec"]["torchTaskSpecs"]
        if task_type not in specs:
            raise KeyError("task type %s is not in spec.torchTaskSpecs" % task_type)
        if num_tasks < 0:
            raise ValueError("numTasks must be >= 0")
        specs[task_type]["numTasks"] = int(num_tasks)
        mm = d["spec"].get("minMembers")
        if isinstance(mm, dict) and task_type in mm and int(mm[task_type]) > num_tasks:
            mm[task_type] = int(num_tasks)     # MinMember may not exceed NumTasks (volcano.go:134-137)
        h = C.c_void_p()
        check(lib().tok_job_parse(json.dumps(d).encode(), C.byref(h)))
        old, self._h = self._h, h
        lib().tok_job_free(old)

    # SetClusterSpec
    def cluster_spec(self, task_type: str, index: int) -> dict:
        return call_json(lib().tok_job_cluster_spec, self._h, task_type.encode(), index)

    def replica_env(self, task_type: str, index: int) -> Dict[str, str]:
        return {e["name"]: e["value"] for e in self.cluster_spec(task_type, index)["env"]}

    # CheckDAGConditionReady
    def dag_ready(self, task_type: str, phases: Dict[str, List[str]]) -> bool:
        r = C.c_int()
        check(lib().tok_job_dag_ready(self._h, task_type.encode(), json.dumps(phases).encode(),
                                      C.byref(r)))
        return bool(r.value)

    # GangScheduler.CreatePodGroup over GPU slots
    def gang_admit(self, free_slots: int) -> dict:
        return call_json(lib().tok_gang_admit, self._h, free_slots)

    # updateJobTaskStatuses + updateGeneralJobStatus
    def update_status(self, replicas: Dict[str, List[dict]], restarting: bool, now: str) -> dict:
        return call_json(lib().tok_job_update_status, self._h, json.dumps(replicas).encode(),
                         int(restarting), now.encode())

    # termination policies of ReconcileJobs (backoffLimit / activeDurations / cleanPodPolicy / TTL)
    def check_termination(self, replicas: Dict[str, List[dict]], prev_retries: int, now: str) -> dict:
        return call_json(lib().tok_job_check_termination, self._h, json.dumps(replicas).encode(),
                         prev_retries, now.encode())

    def set_condition(self, ctype: str, reason: str, message: str, now: str) -> None:
        check(lib().tok_job_set_condition(self._h, ctype.encode(), reason.encode(),
                                          message.encode(), now.encode()))

    def need_enqueue(self) -> bool:
        r = C.c_int()
        check(lib().tok_job_need_enqueue(self._h, C.byref(r)))
        return bool(r.value)

    @property
    def status(self) -> dict:
        return self.to_dict().get("status", {})

    def last_condition(self) -> Optional[str]:
        conds = self.status.get("conditions") or []
        return conds[-1]["type"] if conds else None  # printer column "State", torchjob_types.go:320

    def close(self):
        if self._h:
            lib().tok_job_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def should_failover(restart_policy: str, exit_code: int, reason: str = "") -> bool:
    """shouldPodFailover (controllers/common/failover.go:52-61)."""
    r = C.c_int()
    check(lib().tok_failover_decide(restart_policy.encode(), exit_code, reason.encode(), C.byref(r)))
    return bool(r.value)


def set_feature_gates(**gates: bool) -> None:
    """--feature-gates (pkg/features/features.go): GangScheduling, DAGScheduling, JobCoordinator,
    TorchLocalMasterAddr, HostNetWithHeadlessSvc."""
    bits = {"GangScheduling": 1, "DAGScheduling": 2, "JobCoordinator": 4, "TorchLocalMasterAddr": 8,
            "HostNetWithHeadlessSvc": 16}
    cur = lib().tok_get_feature_gates()
    for k, v in gates.items():
        cur = (cur | bits[k]) if v else (cur & ~bits[k])
    check(lib().tok_set_feature_gates(cur))
