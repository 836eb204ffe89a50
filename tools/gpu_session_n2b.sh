#!/bin/bash
# Second 2-GPU session: re-run the failed parity cases, controller-driven GPU tests, barrier anatomy,
# PDL on/off for the zero-copy exchange.
set -x
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1500 python -m pytest tests/test_allreduce_gpu.py tests/test_controller_gpu.py -m gpu -q -rs -k "zero_copy or golden or nvls or broadcast or dead_peer or elastic or controller or queued or pool" > gpurun_out/n2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2b_pytest.log
timeout 300 $T --master-port 29621 tools/barrier_bench.py > gpurun_out/n2b_barrier.log 2>&1
TOK_SYMM_POOL_MB=2048 TOK_PDL=0 timeout 400 $T --master-port 29622 tools/sweep.py --min-kb 64 --max-mb 64 --dtypes bf16 --algos 3 --zero-copy --no-nccl --iters 30 --out gpurun_out/n2b_sweep_pdl0 > gpurun_out/n2b_sweep_pdl0.log 2>&1
TOK_SYMM_POOL_MB=2048 TOK_PDL=1 timeout 400 $T --master-port 29623 tools/sweep.py --min-kb 64 --max-mb 64 --dtypes bf16 --algos 3 --zero-copy --no-nccl --iters 30 --out gpurun_out/n2b_sweep_pdl1 > gpurun_out/n2b_sweep_pdl1.log 2>&1
tail -n 8 gpurun_out/n2b_pytest.log
cat gpurun_out/n2b_barrier.log | grep "^{"
