#!/bin/bash
# r2 GPU call B (1 GPU): lazy diagonals, light/full kernel variants, linear op walk.  gpurun --timeout 1500 -- 'bash scripts/gpu_r2_b.sh'
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/r2_b.jsonl
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --skip-cpu-baseline "$@" 2>>gpurun_out/r2_b.err | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name', '%s=%.0f ms/step=%.1f launches=%d e2e=%.0f check=%s'%(j['unit'],j['value'],j['ms_per_step'],j['gpu_launches'],j['e2e']['value'],j.get('check'))); j['run']='$name'; open('gpurun_out/r2_b.jsonl','a').write(json.dumps(j)+'\n')"
}
echo "== bench"
run htcnot X=1 -- --steps 5 --warmup 3
run htcnot_nolazy B200SV_LAZY_DIAG=0 -- --steps 3 --warmup 3 --skip-check
run htcnot_forcefull B200SV_FORCE_FULL=1 -- --steps 3 --warmup 3 --skip-check
run htcnot_L5 B200SV_FUSED=4,5,5,7 -- --steps 3 --warmup 3 --skip-check
run htcnot_rb3 B200SV_FUSED=3,6,6,7 -- --steps 3 --warmup 3 --skip-check
run qv X=1 -- --steps 3 --warmup 3 --workload qv --depth 40
run qft64 X=1 -- --steps 5 --warmup 3 --workload qft --precision 64
run qft64_rb3_2cta B200SV_FUSED=4,6,6,7,3,2 -- --steps 5 --warmup 3 --workload qft --precision 64 --skip-check
run qft64_rb3_3cta B200SV_FUSED=4,6,6,7,3,3 -- --steps 5 --warmup 3 --workload qft --precision 64 --skip-check
run qft32 X=1 -- --steps 5 --warmup 3 --workload qft --precision 32
run grover30 X=1 -- --steps 3 --warmup 3 --workload grover --depth 3
echo "== ncu full (28 q)"
NCU_OUT=prof_fused_r2b bash scripts/gpu_ncu_full.sh
echo "== done"
