#!/bin/bash
# First GPU call of the next round (1 GPU): time what r1 could only validate on CPU, then the usual full check.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_r2_first.sh'
set -u
mkdir -p gpurun_out
echo "== planner search on/off, LAYER bundling, tile shape (bench, 2 steps each)"
: > gpurun_out/r2_first.jsonl
for cfg in "search2:B200SV_PLAN_SEARCH=2" "search0:B200SV_PLAN_SEARCH=0" "search4:B200SV_PLAN_SEARCH=4" \
           "search2_bundle3:B200SV_PLAN_SEARCH=2 B200SV_FUSED=4,6,6,3" "search2_L5:B200SV_PLAN_SEARCH=2 B200SV_FUSED=4,5,5,7"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  for wl in htcnot qv; do
    env $envs timeout 600 python bench.py --steps 2 --warmup 3 --skip-cpu-baseline --workload $wl 2>>gpurun_out/r2_first.err | python -c "
import sys,json
for l in sys.stdin:
    j=json.loads(l); print('$name $wl', '%s=%.0f ms/step=%.1f launches=%d e2e=%.0f'%(j['unit'],j['value'],j['ms_per_step'],j['gpu_launches'],j['e2e']['value'])); j['run']='$name $wl'; open('gpurun_out/r2_first.jsonl','a').write(json.dumps(j)+'\n')"
  done
done
echo "== full check"
bash scripts/gpu_check.sh
echo "== reference tests on the drop-in"
bash scripts/gpu_dropin.sh
# then, on 2 and 8 GPUs:  N=2 bash scripts/gpu_multi.sh ; B200SV_SHARD_DEFER=0 N=2 SKIP_TESTS=1 SKIP_NCCL=1 bash scripts/gpu_multi.sh
