#!/bin/bash
# One 2-GPU session: GPU tests (NVLS + multicast barrier + broadcast at N=2), phase stamps, sweep, bench.
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests/test_allreduce_gpu.py -m gpu -x -q -rs -k "not p2p_bit_exact_threads and not world1" > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2_pytest.log
CTAS=64 SIZES_MB=1,4,27,128 timeout 300 $T --master-port 29611 tools/phase_breakdown.py > gpurun_out/n2_phases.log 2>&1
cp gpurun_out/phases_n2.json gpurun_out/n2_phases.json
TOK_SYMM_POOL_MB=3072 timeout 600 $T --master-port 29612 tools/sweep.py --max-mb 256 --dtypes bf16 --zero-copy --iters 30 --out gpurun_out/n2_sweep > gpurun_out/n2_sweep.log 2>&1
timeout 600 $T --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/n2_bench.log
timeout 900 $T --master-port 29614 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > gpurun_out/n2_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/n2_ref.log
tail -n 5 gpurun_out/n2_pytest.log
tail -n 3 gpurun_out/n2_bench.log | cut -c1-400
