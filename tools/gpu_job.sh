#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "reverb or render or cfg3 or cfg5" 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg3.csv python tools/run_workload_once.py cfg3 > /dev/null 2>&1
python tools/launch_agg.py gpurun_out/r02_launches_cfg3.csv 3 | head -7
