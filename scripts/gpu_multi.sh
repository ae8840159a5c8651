#!/bin/bash
# multi-GPU: NCCL parity test + sharded bench at N GPUs (run with gpurun --gpus N)
set -u
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/gpus_$N.txt
if [ "${SKIP_TESTS:-0}" != "1" ]; then
echo "== sharded parity (nccl)"; timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_multi_$N.log
fi
if [ "${SKIP_NCCL:-0}" != "1" ]; then
echo "== bench N=$N (NCCL exchange)"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 --exchange nccl ${BENCH_EXTRA:-} 2>gpurun_out/bench_multi_nccl_$N.err | tee gpurun_out/bench_multi_nccl_$N.json
fi
echo "== bench N=$N"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 ${BENCH_EXTRA:-} 2>gpurun_out/bench_multi_$N.err | tee gpurun_out/bench_multi_$N.json
tail -5 gpurun_out/bench_multi_$N.err
