#!/bin/bash
# one `ncu --set full` capture of the top kernel (1 GPU), report lands in gpurun_out/
set -u
mkdir -p gpurun_out
KREGEX=${KREGEX:-k_fused_sweep}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -s ${NCU_SKIP:-4} -c ${NCU_COUNT:-2} \
    -f -o gpurun_out/${NCU_OUT:-prof_fused} python bench.py --steps 1 --warmup 3 --qubits ${NCU_QUBITS:-28} --skip-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
