#!/bin/bash
# r2 first GPU call (1 GPU): parity of the rewritten schedule + STAGE/PH2 ops, then timings old-vs-new schedule, other workloads,
# one ncu --set full capture.   gpurun --timeout 1500 -- 'bash scripts/gpu_r2_a.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/r2_a.jsonl
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --skip-cpu-baseline "$@" 2>>gpurun_out/r2_a.err | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name', '%s=%.0f ms/step=%.1f launches=%d e2e=%.0f'%(j['unit'],j['value'],j['ms_per_step'],j['gpu_launches'],j['e2e']['value'])); j['run']='$name'; open('gpurun_out/r2_a.jsonl','a').write(json.dumps(j)+'\n')"
}
echo "== bench"
run htcnot_new B200SV_REWRITE=1 -- --steps 5 --warmup 3
run htcnot_norewrite B200SV_REWRITE=0 -- --steps 3 --warmup 3
run htcnot_L5 B200SV_FUSED=4,5,5,7 -- --steps 3 --warmup 3
run htcnot_L7 B200SV_FUSED=4,7,7,7 -- --steps 3 --warmup 3
run qv_new B200SV_REWRITE=1 -- --steps 3 --warmup 3 --workload qv
run qv_norewrite B200SV_REWRITE=0 -- --steps 2 --warmup 3 --workload qv
run qft64_new B200SV_REWRITE=1 -- --steps 5 --warmup 3 --workload qft --precision 64
run qft64_rb4 B200SV_FUSED=4,6,6,7,4 -- --steps 5 --warmup 3 --workload qft --precision 64
run qft32_new B200SV_REWRITE=1 -- --steps 5 --warmup 3 --workload qft --precision 32
echo "== ncu full (28 q)"
NCU_OUT=prof_fused_r2a bash scripts/gpu_ncu_full.sh
echo "== dropin (reference tests on the C++ adapter)"
bash scripts/gpu_dropin.sh 2>&1 | tail -40
echo "== done"
