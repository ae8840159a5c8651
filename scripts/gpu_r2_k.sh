#!/bin/bash
# r2 last GPU call (1 GPU, ~1 minute): smoke() + the fastest parity tests on the final build
set -u
mkdir -p gpurun_out
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
timeout 45 python -m pytest tests/test_exchange_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "exchange or golden" 2>&1 | tail -4 | tee gpurun_out/final_quick_tests.log
