#!/bin/bash
# r2 GPU call F (1 GPU): three CTAs per SM for small sweep programs (B200SV_MINB3) A/B on one box, then the full gpu suite
set -u
mkdir -p gpurun_out
: > gpurun_out/r2_f.jsonl
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --skip-cpu-baseline "$@" 2>>gpurun_out/r2_f.err | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name', '%s=%.0f ms/step=%.1f launches=%d e2e=%.0f check=%s'%(j['unit'],j['value'],j['ms_per_step'],j['gpu_launches'],j['e2e']['value'],(j.get('check') or {}).get('ok'))); j['run']='$name'; open('gpurun_out/r2_f.jsonl','a').write(json.dumps(j)+'\n')"
}
echo "== A/B on one box"
run htcnot_2cta X=1 -- --steps 4 --warmup 3 --skip-check
run htcnot_3cta B200SV_MINB3=1 -- --steps 4 --warmup 3
run htcnot_2cta_again X=1 -- --steps 4 --warmup 3 --skip-check
run qv_2cta X=1 -- --steps 3 --warmup 3 --workload qv --depth 40 --skip-check
run qv_3cta B200SV_MINB3=1 -- --steps 3 --warmup 3 --workload qv --depth 40
run qft32_2cta X=1 -- --steps 5 --warmup 3 --workload qft --precision 32 --skip-check
run qft32_3cta B200SV_MINB3=1 -- --steps 5 --warmup 3 --workload qft --precision 32
run grover_2cta X=1 -- --steps 3 --warmup 3 --workload grover --depth 3 --skip-check
run grover_3cta B200SV_MINB3=1 -- --steps 3 --warmup 3 --workload grover --depth 3
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
echo "== done"
