#!/bin/bash
# Second 8-GPU session: config 3 of BASELINE (two queued jobs) with the corrected submission order,
# grid / unroll tuning of the zero-copy NVLS kernel with the acquire-free trailing barrier, bench with
# the best setting.
set -x
mkdir -p gpurun_out
export TOK_BARRIER_TIMEOUT_MS=45000
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
OUT=zc_tune2_n8.json CTAS=64,96,128 UNROLL=8,16 timeout 200 $T8 --master-port 29801 tools/zc_tune.py > gpurun_out/n8b_zc_tune.log 2>&1
cat gpurun_out/zc_best.env; source gpurun_out/zc_best.env
timeout 300 $T8 --master-port 29802 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/n8b_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/n8b_bench.log
unset TOK_ZC_CTAS TOK_NVLS_UNROLL
timeout 300 python tools/run_cfg4.py --out gpurun_out/n8b_cfg4.json --resnet-steps 120 --bert-steps 200 > gpurun_out/n8b_cfg4.log 2>&1; echo "cfg4 rc=$?" >> gpurun_out/n8b_cfg4.log
grep "^{" gpurun_out/n8b_zc_tune.log; tail -n 2 gpurun_out/n8b_bench.log | cut -c1-200
