#!/bin/bash
# One 1-GPU session: GPU test suite, smoke, local-kernel variants, bench, launch list, ncu capture.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/n1_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/n1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n1_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n1_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/n1_smoke.log
timeout 600 python tools/local_bench.py --out gpurun_out/n1_local_bench.json > gpurun_out/n1_local_bench.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/n1_bench.log 2>&1; echo "bench rc=$?" >> gpurun_out/n1_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/n1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/n1_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:local -c 12 -o gpurun_out/n1_local_prof -f python tools/local_bench.py --ncu > gpurun_out/n1_ncu_local.log 2>&1
tail -5 gpurun_out/n1_pytest.log gpurun_out/n1_smoke.log gpurun_out/n1_bench.log
