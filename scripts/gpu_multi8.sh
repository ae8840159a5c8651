#!/bin/bash
# Multi-GPU call (gpurun --gpus N, N = 2 or 8): BASELINE configs[3] (QV, sharded engine AND the unchanged reference QPager over
# the drop-in, one page per GPU) and configs[4] (Grover), the headline circuit, and the sharded parity tests.
set -u
N=${N:-8}
K=$(python -c "print(($N).bit_length()-1)")
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv | tee gpurun_out/gpus_$N.txt
export LD_LIBRARY_PATH=$PWD/qrack_b200:${LD_LIBRARY_PATH:-}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== sharded parity tests (nccl + p2p kernel, 26 q vs compiled reference)"
timeout 1500 python -m pytest tests/test_sharded_gpu.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_multi_$N.log
fi
: > gpurun_out/multi_$N.jsonl
bench() { # name, args...
  local name=$1; shift
  timeout 1500 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" 2>gpurun_out/bench_${name}_$N.err | tee gpurun_out/bench_${name}_$N.json | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name N=$N', '%s=%.0f ms/step=%.1f e2e=%.0f'%(j['unit'],j['value'],j['ms_per_step'],j['e2e']['value']), 'sharding=',j.get('sharding'), 'check=',j.get('check')); j['run']='$name'; open('gpurun_out/multi_$N.jsonl','a').write(json.dumps(j)+'\n')"
  tail -3 gpurun_out/bench_${name}_$N.err
}
echo "== bench htcnot (headline circuit, 30+$K qubits)"; bench htcnot --steps ${STEPS:-3} --warmup 3
echo "== NVLink counter check: Tx bytes of GPU 0 over a run of 7 circuit replays (3 warm-up + 2 timed + 2 e2e, no mirror check)"
nvl() { nvidia-smi nvlink -gt d -i 0 2>/dev/null | awk '/Data Tx/ {s += $5} END {printf "%.0f\n", s * 1024}'; }
T0=$(nvl)
timeout 900 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 2 --warmup 3 --skip-check 2>/dev/null > gpurun_out/bench_nvl_$N.json
T1=$(nvl)
python - <<PY | tee gpurun_out/nvlink_check_$N.txt
import json
j = json.loads(open('gpurun_out/bench_nvl_$N.json').read().strip().splitlines()[-1])
model = j['sharding']['nvlink_bytes_out_per_gpu_per_step'] * 7
meas = $T1 - $T0
print('NVLink Tx of GPU 0 over 7 replays: counter delta %.2f GB, model (exchanges x page x (N-1)/N) %.2f GB, ratio %.3f' % (meas / 1e9, model / 1e9, meas / max(model, 1)))
PY
echo "== bench qv (configs[3]: 30+$K qubits, depth = qubits)"; bench qv --steps ${STEPS:-2} --warmup 3 --workload qv
echo "== bench grover (configs[4]: 31+$K qubits)"; bench grover --steps ${STEPS:-2} --warmup 3 --workload grover --qubits 31 --depth 3
echo "== the unchanged reference QPager over the drop-in, one page per GPU (QRACK_QPAGER_DEVICES)"
DEVS=$(python -c "print(','.join(str(i) for i in range($N)))")
python - <<PY
import sys; sys.path.insert(0,'.')
from qrack_b200 import qscript
n = 30 + $K
open('/tmp/qv_big.qs','w').write(qscript.quantum_volume(n, depth=n, seed=33, timed=True) + "Norm\n" + "".join("Prob %d\n" % q for q in (0, 7, n - 2, n - 1)))
m = 20 + $K
t = qscript.quantum_volume(m, depth=6, seed=5, timed=False)
open('/tmp/qv_small.qs','w').write(t)
PY
echo "-- parity at 20+$K qubits: pager-cuda:20 over $N devices vs reference QEngineCPU"
timeout 600 oracle/_ref/ref_harness_f32 /tmp/qv_small.qs --dump /tmp/ref_small >/dev/null 2>&1
QRACK_QPAGER_DEVICES=$DEVS timeout 600 dropin/_build/harness_b200_f32 /tmp/qv_small.qs --dump /tmp/dev_small --engine pager-cuda:20 2>&1 | tail -2
python - <<'PY' | tee gpurun_out/qpager_multi_parity.log
import numpy as np
a=np.fromfile('/tmp/ref_small.0.bin',dtype=np.complex64); b=np.fromfile('/tmp/dev_small.0.bin',dtype=np.complex64)
print('QPager over drop-in, pages on several GPUs: max |delta amp| vs QEngineCPU = %.3e (n=%d)' % (np.abs(a-b).max(), int(np.log2(a.size))))
PY
echo "-- timing: QV 30+$K qubits, pager-cuda:30 (one 2^30 page per GPU)"
QRACK_QPAGER_DEVICES=$DEVS timeout 1500 dropin/_build/harness_b200_f32 /tmp/qv_big.qs --engine pager-cuda:30 --time --results gpurun_out/qpager_qv_results_$N.txt 2>&1 | tail -3 | tee gpurun_out/qpager_qv_$N.log
cat gpurun_out/qpager_qv_results_$N.txt 2>/dev/null | head
nvidia-smi nvlink -gt d 2>/dev/null | head -20 > gpurun_out/nvlink_counters_$N.txt
echo "== done"
