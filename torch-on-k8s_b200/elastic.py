"""torchelastic replica-count policy over the C ABI (csrc/ctl_elastic.cpp)."""
from __future__ import annotations

import ctypes as C

from ._ffi import call_json, check, lib
from .job import TorchJob

LOOP_PERIOD_S = 30.0   # elastictorchjob_controller.go:59-61
METRIC_COUNT = 5


def parse_log_line(line: str) -> dict:
    """getMetricsObservation's parser (observation.go:54-76)."""
    return call_json(lib().tok_elastic_parse_log, line.encode())


class ElasticPolicy:
    def __init__(self, metric_count: int = METRIC_COUNT):
        self._h = C.c_void_p()
        check(lib().tok_elastic_create(metric_count, C.byref(self._h)))

    def observe(self, job: TorchJob, latency: float, *, has_pending: bool = False,
                has_failed: bool = False) -> dict:
        """One decision pass; on scale/revert job.spec.torchTaskSpecs.Worker.numTasks is updated."""
        return call_json(lib().tok_elastic_observe, self._h, job._h, C.c_double(latency),
                         int(has_pending), int(has_failed))

    def close(self):
        if self._h:
            lib().tok_elastic_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
