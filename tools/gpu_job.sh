#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "iir or cascade or cfg2 or kat or low_frequency or geq or next_geq or backward or render" 2>&1 | grep -E "passed|failed|FAILED|^E " | head -10 > gpurun_out/r02_casc.log
timeout 300 python tools/quick_time.py 2>&1 | head -5 >> gpurun_out/r02_casc.log
cat gpurun_out/r02_casc.log
