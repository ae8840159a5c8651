#!/bin/bash
# r2 GPU call I (1 GPU): the single-device re-page tests and the TMA staging probe
set -u
mkdir -p gpurun_out
echo "== re-page primitives on one device"
timeout 600 python -m pytest tests/test_exchange_gpu.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_exchange_gpu.log
echo "== TMA staging probe"
timeout 120 scripts/probes/tma_stage_probe 30 5 | tee gpurun_out/tma_stage_probe.jsonl
echo "== done"
