#!/bin/bash
# r2 GPU call J (gpurun --gpus N; the run of r2 had VIRT on by default, it is opt-in now): tail carry across exchanges (b200sv_flush_carry + rank bits as virtual qubits) on/off, same box, back to back,
# after the sharded parity tests (which run with the defaults: pull-mode exchange, carry on)
set -u
N=${N:-2}
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/qrack_b200:${LD_LIBRARY_PATH:-}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== sharded parity tests"
timeout 1500 python -m pytest tests/test_sharded_gpu.py -q -m gpu ${TESTSEL:+-k "$TESTSEL"} 2>&1 | tail -8 | tee gpurun_out/pytest_carry_$N.log
fi
: > gpurun_out/carry_$N.jsonl
bench() { # name, env, args...
  local name=$1; shift; local ev=$1; shift
  env $ev timeout 900 $TR --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" 2>gpurun_out/bench_${name}_$N.err | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$name N=$N', '%s=%.0f ms/step=%.1f e2e=%.0f'%(j['unit'],j['value'],j['ms_per_step'],j['e2e']['value']), 'sharding=',j.get('sharding'), 'check=',(j.get('check') or {}).get('ok')); j['run']='$name'; open('gpurun_out/carry_$N.jsonl','a').write(json.dumps(j)+'\n')"
  tail -2 gpurun_out/bench_${name}_$N.err
}
bench htcnot_carry "B200SV_SHARD_VIRT=1 B200SV_SHARD_CARRY=1000000" --steps 3 --warmup 3
bench htcnot_nocarry "B200SV_SHARD_VIRT=1 B200SV_SHARD_CARRY=0" --steps 3 --warmup 3 --skip-check
bench qv_carry "B200SV_SHARD_VIRT=1 B200SV_SHARD_CARRY=1000000" --steps 2 --warmup 3 --workload qv
[ "${QVOFF:-1}" = "1" ] && bench qv_nocarry "B200SV_SHARD_VIRT=1 B200SV_SHARD_CARRY=0" --steps 2 --warmup 3 --workload qv --skip-check
echo "== done"
