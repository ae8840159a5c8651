#!/bin/bash
# r2 GPU call H (gpurun --gpus N): the exchange in PULL mode (b200sv_exchange_pull: the re-page rides on the first fused sweep of the next
# window) against the push kernel (B200SV_SHARD_PULL=0), same box, back to back; before that the sharded parity tests in all three modes.
set -u
N=${N:-2}
K=$(python -c "print(($N).bit_length()-1)")
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/qrack_b200:${LD_LIBRARY_PATH:-}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== sharded parity tests (nccl / push kernel / pull fused into the sweep; 26 q vs compiled reference)"
timeout 1500 python -m pytest tests/test_sharded_gpu.py -q -m gpu ${TESTSEL:+-k "$TESTSEL"} 2>&1 | tail -8 | tee gpurun_out/pytest_pull_$N.log
fi
: > gpurun_out/pull_$N.jsonl
bench() { # name, env, args...
  local name=$1; shift; local ev=$1; shift
  env $ev timeout 900 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" 2>gpurun_out/bench_${name}_$N.err | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name N=$N', '%s=%.0f ms/step=%.1f e2e=%.0f'%(j['unit'],j['value'],j['ms_per_step'],j['e2e']['value']), 'sharding=',j.get('sharding'), 'check=',(j.get('check') or {}).get('ok')); j['run']='$name'; open('gpurun_out/pull_$N.jsonl','a').write(json.dumps(j)+'\n')"
  tail -2 gpurun_out/bench_${name}_$N.err
}
bench htcnot_pull B200SV_SHARD_PULL=1 --steps 3 --warmup 3
[ "${PUSH:-1}" = "1" ] && bench htcnot_push B200SV_SHARD_PULL=0 --steps 3 --warmup 3 --skip-check
if [ "${QV:-1}" = "1" ]; then
bench qv_pull B200SV_SHARD_PULL=1 --steps 2 --warmup 3 --workload qv
[ "${PUSH:-1}" = "1" ] && bench qv_push B200SV_SHARD_PULL=0 --steps 2 --warmup 3 --workload qv --skip-check
fi
if [ "${GROVER:-1}" = "1" ]; then
bench grover_pull B200SV_SHARD_PULL=1 --steps 2 --warmup 3 --workload grover --qubits 31 --depth 3
bench grover_push B200SV_SHARD_PULL=0 --steps 2 --warmup 3 --workload grover --qubits 31 --depth 3 --skip-check
fi
echo "== done"
