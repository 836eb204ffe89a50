"""kubectl-shaped views over the controller's state directory:

    python -m torch_on_k8s_b200 run job.yaml [...] --state-dir state/     # the controller
    python -m torch_on_k8s_b200 get torchjob [-n NS] --state-dir state/
    python -m torch_on_k8s_b200 describe torchjob NAME [-n NS] --state-dir state/

`get` prints the reference's printer columns (apis/train/v1alpha1/torchjob_types.go:320-324): State =
last condition, Age, Model-Version, Max-Lifetime, TTL-After-Finished.  (The reference's JSONPaths
`.spec.activeDeadlineSeconds` / `.spec.ttlSecondsAfterFinished` do not exist in its own spec — the
fields are `activeDurations` / `TTLSecondsAfterFinished`; the real fields are shown here.)
"""
from __future__ import annotations

import calendar
import json
import os
import sys
import time
from typing import List


def _age(ts: str) -> str:
    try:
        t = calendar.timegm(time.strptime(ts[:19], "%Y-%m-%dT%H:%M:%S"))
    except Exception:  # noqa: BLE001
        return "<unknown>"
    s = max(0, int(time.time() - t))
    for unit, n in (("d", 86400), ("h", 3600), ("m", 60)):
        if s >= n:
            return "%d%s" % (s // n, unit)
    return "%ds" % s


def load_jobs(state_dir: str, namespace: str = "") -> List[dict]:
    jobs = []
    if not os.path.isdir(state_dir):
        return jobs
    for f in sorted(os.listdir(state_dir)):
        if f.endswith(".json"):
            with open(os.path.join(state_dir, f)) as fh:
                d = json.load(fh)
            if not namespace or d["metadata"].get("namespace", "default") == namespace:
                jobs.append(d)
    return jobs


def cmd_get(state_dir: str, namespace: str = "", out=sys.stdout) -> int:
    rows = [("NAME", "STATE", "AGE", "MODEL-VERSION", "MAX-LIFETIME", "TTL-AFTER-FINISHED")]
    for d in load_jobs(state_dir, namespace):
        conds = (d.get("status") or {}).get("conditions") or []
        rows.append((d["metadata"]["name"], conds[-1]["type"] if conds else "",
                     _age(d["metadata"].get("creationTimestamp", "")),
                     (d.get("status") or {}).get("modelVersionName", ""),
                     str(d["spec"].get("activeDurations", "")),
                     str(d["spec"].get("TTLSecondsAfterFinished", ""))))
    widths = [max(len(r[i]) for r in rows) for i in range(len(rows[0]))]
    for r in rows:
        out.write("   ".join(c.ljust(w) for c, w in zip(r, widths)).rstrip() + "\n")
    return 0


def cmd_describe(state_dir: str, name: str, namespace: str = "", out=sys.stdout) -> int:
    for d in load_jobs(state_dir, namespace):
        if d["metadata"]["name"] == name:
            st = d.get("status") or {}
            out.write("Name:         %s\nNamespace:    %s\nAPI Version:  %s\nKind:         %s\n" %
                      (name, d["metadata"].get("namespace", "default"), d.get("apiVersion", ""),
                       d.get("kind", "")))
            out.write("Task Statuses:\n")
            for tt, t in (st.get("taskStatuses") or {}).items():
                out.write("  %-9s active=%d succeed=%d failed=%d\n" %
                          (tt, t.get("active", 0), t.get("succeed", 0), t.get("failed", 0)))
            out.write("Conditions:\n")
            for c in st.get("conditions") or []:
                out.write("  %-10s %-5s %-14s %s  %s\n" % (c["type"], c["status"], c.get("reason", ""),
                                                         c.get("lastTransitionTime", ""),
                                                         c.get("message", "")))
            es = st.get("elasticScalingStatues") or {}
            for tt, e in es.items():
                out.write("Elastic (%s): %s\n" % (tt, json.dumps(e)))
            out.write("Events:\n")
            for ev in d.get("x-events", []):
                out.write("  %s  %s  %s\n" % tuple(ev[:3]))
            return 0
    out.write('Error from server (NotFound): torchjobs.train.distributed.io "%s" not found\n' % name)
    return 1


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0
    verb = argv.pop(0)
    if verb == "run":
        from .controller import main as run
        return run(argv)

    def opt(flag, default=""):
        if flag in argv:
            i = argv.index(flag)
            v = argv[i + 1]
            del argv[i:i + 2]
            return v
        return default
    state = opt("--state-dir", os.environ.get("TOK8S_STATE_DIR", "tok8s-state"))
    ns = opt("-n")
    if argv and argv[0] in ("torchjob", "torchjobs", "tj"):
        argv.pop(0)
    if verb == "get":
        return cmd_get(state, ns)
    if verb == "describe" and argv:
        return cmd_describe(state, argv[0], ns)
    print("usage: python -m torch_on_k8s_b200 {run|get|describe} ...", file=sys.stderr)
    return 2
