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
echo "== kernel to beat: the reference's own QEngineCUDA (oracle/_ref, sm_100) vs the drop-in, same harness, same script, same box"
python - <<'PY'
import sys; sys.path.insert(0,'.')
from qrack_b200 import qscript
open('/tmp/c2.qs','w').write(qscript.random_htcnot(30,40,seed=20250921,timed=True))
open('/tmp/c2s.qs','w').write(qscript.random_htcnot(28,40,seed=20250921,timed=True))
PY
export LD_LIBRARY_PATH=$PWD/qrack_b200:${LD_LIBRARY_PATH:-}
for sc in c2s c2; do
  echo "-- $sc reference QEngineCUDA"; timeout 600 oracle/_ref/ref_harness_cuda_f32 /tmp/$sc.qs --engine cuda --time 2>&1 | tail -2 | tee -a gpurun_out/refcuda_harness.log
  echo "-- $sc drop-in QEngineCUDA";   timeout 600 dropin/_build/harness_b200_f32 /tmp/$sc.qs --engine cuda --time 2>&1 | tail -2 | tee -a gpurun_out/refcuda_harness.log
done
for t in test_qft_permutation_init test_random_circuit_sampling; do
  echo "-- benchmarks $t: reference CUDA engine"; timeout 900 oracle/_ref/benchmarks_refcuda --layer-qengine --proc-cuda --single -m 30 --samples 3 --disable-terminal-measurement --disable-hardware-rng $t 2>&1 | tail -12 | tee -a gpurun_out/refcuda_bench.log
  echo "-- benchmarks $t: drop-in"; timeout 900 dropin/_build/f32/benchmarks_b200 --layer-qengine --proc-cuda --single -m 30 --samples 3 --disable-terminal-measurement --disable-hardware-rng $t 2>&1 | tail -12 | tee -a gpurun_out/refcuda_bench.log
done
echo "== ncu full (28 q)"
NCU_OUT=prof_fused_r2b bash scripts/gpu_ncu_full.sh
echo "== done"
