"""`bench.py --impl reference` must finish when it is launched the way the driver launches it for
N > 1 — under `torch.distributed.run`, whose environment (TORCHELASTIC_USE_AGENT_STORE, LOCAL_RANK,
OMP_NUM_THREADS=1, ...) round 1's arm leaked into its gloo children, which then waited for an agent
store nobody had started.  Replicas get a container-clean environment: the SetClusterSpec variables
(controllers/train/torchjob_controller.go:394-446) and nothing of the launcher."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_clean_env_drops_launcher_variables(monkeypatch):
    from oracle.gloo_torchjob import clean_env, replica_env
    for k, v in dict(TORCHELASTIC_USE_AGENT_STORE="True", TORCHELASTIC_RUN_ID="x", LOCAL_RANK="3",
                     GROUP_RANK="0", ROLE_RANK="3", OMP_NUM_THREADS="1", RANK="3", WORLD_SIZE="8",
                     MASTER_PORT="1", NCCL_DEBUG="INFO", TOK8S_JOB="j", KEEP_ME="yes").items():
        monkeypatch.setenv(k, v)
    env = clean_env(replica_env("job", "worker", 0, 1, port=23456))
    assert env["KEEP_ME"] == "yes" and env["RANK"] == "1" and env["WORLD_SIZE"] == "2"
    assert env["MASTER_PORT"] == "23456" and env["PYTHONUNBUFFERED"] == "0"
    for k in ("TORCHELASTIC_USE_AGENT_STORE", "TORCHELASTIC_RUN_ID", "LOCAL_RANK", "GROUP_RANK",
              "ROLE_RANK", "OMP_NUM_THREADS", "NCCL_DEBUG", "TOK8S_JOB"):
        assert k not in env


def test_reference_arm_finishes_under_torchrun_world2():
    env = dict(os.environ)
    env.pop("PYTEST_CURRENT_TEST", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29591", os.path.join(ROOT, "bench.py"),
           "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--ref-batch", "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                       # rank 0 alone prints; rank 1 exits without work
    line = lines[0]
    sys.path.insert(0, ROOT)
    import bench
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["config"] == bench.workload_config(256, 2)          # the SAME config object as our arm
    assert line["cpu_baseline"]["kind"] == "reference" and line["sample_batch_per_replica"] == 2
    assert line["e2e"] == {"value": line["value"], "unit": "images/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
