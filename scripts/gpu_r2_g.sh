#!/bin/bash
# r2 final GPU call (1 GPU): the bench contract lines of both arms, the ncu launch list of the same command, one ncu --set full capture
set -u
mkdir -p gpurun_out
echo "== re-page primitives on one device (pull fused into the sweep / gather / push)"
timeout 600 python -m pytest tests/test_exchange_gpu.py -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_exchange_gpu.log
echo "== bench (default flags, incl. cpu_baseline)"
timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench.err | tee gpurun_out/bench.json
echo "== bench --impl reference"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_reference.json
echo "== configs[2]: QFT fp64"
timeout 600 python bench.py --steps 5 --warmup 3 --workload qft --precision 64 --skip-cpu-baseline 2>>gpurun_out/bench.err | tee gpurun_out/bench_qft64.json
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu-baseline --skip-check > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full (28 q)"
NCU_OUT=prof_fused_r2_final bash scripts/gpu_ncu_full.sh
echo "== done"
