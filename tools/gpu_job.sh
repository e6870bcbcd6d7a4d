#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x -k "dynamics or cfg4 or envelope or approx or ballistics or slow_pole or compressor or noisegate or render or captured" 2>&1 | tail -6 > gpurun_out/r02_dyn_tests.log
timeout 200 python tools/dyn_time.py >> gpurun_out/r02_dyn_tests.log 2>&1
cat gpurun_out/r02_dyn_tests.log
