"""Job coordinator over the C ABI (csrc/ctl_coord.cpp): pkg/coordinator's tenant queues, RR / WRR
queue selection, Quota filter (GPU slots of the box instead of ResourceQuota), Priority score."""
from __future__ import annotations

import ctypes as C
from typing import Optional

from . import _ffi
from ._ffi import call_json, check, lib
from .job import TorchJob

SCHEDULING_PERIOD_S = 0.1  # plugins/registry.go:27


class Coordinator:
    def __init__(self, policy: str = "wrr", weight_mode: str = "replicas", seed: int = 1):
        self._h = C.c_void_p()
        check(lib().tok_coord_create(_ffi.TOK_POLICY_WRR if policy == "wrr" else _ffi.TOK_POLICY_RR,
                                     _ffi.TOK_WRR_WEIGHT_TASK_TYPES if weight_mode == "task_types"
                                     else _ffi.TOK_WRR_WEIGHT_REPLICAS, seed, C.byref(self._h)))

    def set_quota(self, tenant: str, hard_slots: int) -> None:
        check(lib().tok_coord_set_quota(self._h, tenant.encode(), hard_slots))

    def set_used(self, tenant: str, used_slots: int) -> None:
        check(lib().tok_coord_set_used(self._h, tenant.encode(), used_slots))

    def enqueue(self, job: TorchJob, uid: str) -> None:  # EnqueueOrUpdate
        check(lib().tok_coord_enqueue(self._h, job._h, uid.encode()))

    def is_queuing(self, uid: str) -> bool:
        r = C.c_int()
        check(lib().tok_coord_is_queuing(self._h, uid.encode(), C.byref(r)))
        return bool(r.value)

    def dequeue(self, uid: str) -> None:
        check(lib().tok_coord_dequeue(self._h, uid.encode()))

    def job_settled(self, uid: str) -> None:
        check(lib().tok_coord_job_settled(self._h, uid.encode()))

    def pending(self, tenant: Optional[str] = None) -> int:
        r = C.c_int()
        check(lib().tok_coord_pending(self._h, (tenant or "").encode(), C.byref(r)))
        return r.value

    def tick(self, now: float) -> dict:  # one schedule() cycle
        return call_json(lib().tok_coord_tick, self._h, C.c_double(now))

    def close(self):
        if self._h:
            lib().tok_coord_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
