#!/bin/bash
# r2 GPU call C (1 GPU): full gpu test suite (incl. drop-in + full-size parity vs the compiled reference), streaming-kernel table
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== streaming kernel table (30 q fp32, 29 q fp64)"
PRECS=32,64 timeout 900 python scripts/gpu_stream_table.py 2>&1 | tail -60 | tee gpurun_out/stream_table.log
echo "== done"
