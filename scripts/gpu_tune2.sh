#!/bin/bash
# parity (fast subset) + bench over (knobs, workload, precision) triples:  TUNE_RUNS="knobs:workload:precision ..."
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "${PYTEST_K:-golden or c1_20q_vs_oracle or positions or families}" 2>&1 | tail -8 | tee gpurun_out/pytest_tune.log
: > gpurun_out/tune2.jsonl
for run in ${TUNE_RUNS:-4,6,6,3:htcnot:32}; do
  IFS=: read -r cfg wl prec <<< "$run"
  B200SV_FUSED=$cfg timeout 600 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --workload $wl --precision $prec ${BENCH_EXTRA:-} 2>>gpurun_out/tune.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('$run', '%s=%.0f ms/step=%.1f launches=%d phys_frac=%.3f e2e=%.0f sm_mhz=%s'%(j['unit'],j['value'],j['ms_per_step'],j['gpu_launches'],j['roofline']['frac'],j['e2e']['value'],j['clocks']['sm_mhz'])); j['run']='$run'; open('gpurun_out/tune2.jsonl','a').write(json.dumps(j)+'\n')"
done
tail -3 gpurun_out/tune.err
