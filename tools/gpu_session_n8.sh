#!/bin/bash
# The 8-GPU session (8x cost): most valuable first, every step under its own timeout.
set -x
mkdir -p gpurun_out
export TOK_BARRIER_TIMEOUT_MS=45000
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/n8_smi.txt
timeout 420 $T8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/n8_bench.log
timeout 600 python -m pytest tests/test_allreduce_gpu.py -m gpu -q -rs -k "nvls or zero_copy or golden or broadcast or exact_patterns or elastic_reform or large_bucket" > gpurun_out/n8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n8_pytest.log
timeout 600 python tools/run_cfg3.py --out gpurun_out/n8_cfg3.json > gpurun_out/n8_cfg3.log 2>&1; echo "cfg3 rc=$?" >> gpurun_out/n8_cfg3.log
timeout 500 python tools/run_cfg4.py --out gpurun_out/n8_cfg4.json --resnet-steps 120 --bert-steps 200 > gpurun_out/n8_cfg4.log 2>&1; echo "cfg4 rc=$?" >> gpurun_out/n8_cfg4.log
TOK_SYMM_POOL_MB=6144 timeout 420 $T8 --master-port 29702 tools/sweep.py --max-mb 1024 --dtypes bf16,f32 --algos 0 --zero-copy --iters 15 --out gpurun_out/n8_sweep > gpurun_out/n8_sweep.log 2>&1
CTAS=16,32,64 timeout 200 $T8 --master-port 29703 tools/zc_tune.py > gpurun_out/n8_zc_tune.log 2>&1
timeout 300 $T4 --master-port 29704 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/n4_bench.log
CTAS=64 SIZES_MB=22,27 timeout 120 $T8 --master-port 29705 tools/phase_breakdown.py > gpurun_out/n8_phases.log 2>&1; cp gpurun_out/phases_n8.json gpurun_out/n8_phases.json
timeout 100 $T8 --master-port 29706 tools/barrier_bench.py > gpurun_out/n8_barrier.log 2>&1
tail -n 3 gpurun_out/n8_pytest.log; tail -n 2 gpurun_out/n8_bench.log | cut -c1-300; tail -n 2 gpurun_out/n8_cfg3.log | cut -c1-300; tail -n 2 gpurun_out/n8_cfg4.log | cut -c1-300
