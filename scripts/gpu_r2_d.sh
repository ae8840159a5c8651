#!/bin/bash
# r2 GPU call D (1 GPU): validate RB=5 sub-blocks, rotation stages, vectorised streaming kernels, single-pass decompose
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/pytest_gpu.log
: > gpurun_out/r2_d.jsonl
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --skip-cpu-baseline "$@" 2>>gpurun_out/r2_d.err | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name', '%s=%.0f ms/step=%.1f launches=%d e2e=%.0f batched=%s check=%s'%(j['unit'],j['value'],j['ms_per_step'],j['gpu_launches'],j['e2e']['value'],(j.get('e2e_batched') or {}).get('value'),(j.get('check') or {}).get('ok'))); j['run']='$name'; open('gpurun_out/r2_d.jsonl','a').write(json.dumps(j)+'\n')"
}
echo "== bench"
run htcnot X=1 -- --steps 5 --warmup 3
run htcnot_rb4 B200SV_RB5=0 -- --steps 3 --warmup 3 --skip-check
run qv X=1 -- --steps 3 --warmup 3 --workload qv --depth 40
run qv_norot B200SV_ROT=0 -- --steps 3 --warmup 3 --workload qv --depth 40 --skip-check
run qv_rot_rb4 B200SV_RB5=0 -- --steps 3 --warmup 3 --workload qv --depth 40 --skip-check
run qft64 X=1 -- --steps 5 --warmup 3 --workload qft --precision 64
run qft64_rb4 B200SV_RB5=0 -- --steps 5 --warmup 3 --workload qft --precision 64 --skip-check
run qft32 X=1 -- --steps 5 --warmup 3 --workload qft --precision 32
run grover30 X=1 -- --steps 3 --warmup 3 --workload grover --depth 3
echo "== streaming kernel table (30 q fp32, 29 q fp64)"
PRECS=32,64 timeout 900 python scripts/gpu_stream_table.py 2>&1 | tail -60 | tee gpurun_out/stream_table.log
echo "== ncu full (28 q)"
NCU_OUT=prof_fused_r2d bash scripts/gpu_ncu_full.sh
echo "== done"
