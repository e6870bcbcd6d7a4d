#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
rm -f gpurun_out/r02_conv_time.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "fir or reverb or render or conv" 2>&1 | tail -6 > gpurun_out/r02_t2.log
timeout 300 python tools/conv_time.py >> gpurun_out/r02_conv_time.log 2>&1
cat gpurun_out/r02_t2.log gpurun_out/r02_conv_time.log
