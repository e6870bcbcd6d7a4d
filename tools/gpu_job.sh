#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "iir or cascade or cfg2 or kat or low_frequency or geq or next_ or backward or render" 2>&1 | tail -6 > gpurun_out/r02_casc.log
timeout 300 python tools/quick_time.py 2>&1 | head -5 >> gpurun_out/r02_casc.log
cat gpurun_out/r02_casc.log
