#!/bin/bash
# r2 GPU call E (1 GPU): same-box A/B of sweep-kernel variants (scripts/build_variants.sh), then the workloads on the default build
set -u
mkdir -p gpurun_out
: > gpurun_out/r2_e.jsonl
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --skip-cpu-baseline "$@" 2>>gpurun_out/r2_e.err | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name', '%s=%.0f ms/step=%.1f launches=%d e2e=%.0f check=%s'%(j['unit'],j['value'],j['ms_per_step'],j['gpu_launches'],j['e2e']['value'],(j.get('check') or {}).get('ok'))); j['run']='$name'; open('gpurun_out/r2_e.jsonl','a').write(json.dumps(j)+'\n')"
}
V=$PWD/qrack_b200/variants
echo "== htcnot 30 q: variants on the same box"
run default X=1 -- --steps 4 --warmup 3 --skip-check
run callB_kernel B200SV_LIB=$V/libb200sv_callB.so -- --steps 4 --warmup 3 --skip-check
run noneg B200SV_LIB=$V/libb200sv_noneg.so -- --steps 4 --warmup 3 --skip-check
run noneg_nopf B200SV_LIB=$V/libb200sv_noneg_nopf.so -- --steps 4 --warmup 3 --skip-check
run noneg_nofast B200SV_LIB=$V/libb200sv_noneg_nofast.so -- --steps 4 --warmup 3 --skip-check
run nopf B200SV_LIB=$V/libb200sv_nopf.so -- --steps 4 --warmup 3 --skip-check
run default_again X=1 -- --steps 4 --warmup 3 --skip-check
echo "== other workloads, default build"
run qv X=1 -- --steps 3 --warmup 3 --workload qv --depth 40
run qv_callB B200SV_LIB=$V/libb200sv_callB.so -- --steps 3 --warmup 3 --workload qv --depth 40 --skip-check
run qft64 X=1 -- --steps 5 --warmup 3 --workload qft --precision 64
run qft64_callB B200SV_LIB=$V/libb200sv_callB.so -- --steps 5 --warmup 3 --workload qft --precision 64 --skip-check
run qft32 X=1 -- --steps 5 --warmup 3 --workload qft --precision 32
run grover30 X=1 -- --steps 3 --warmup 3 --workload grover --depth 3
echo "== parity smoke of the default build"; timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "knobs or golden or c1_20q or families" 2>&1 | tail -4
echo "== done"
