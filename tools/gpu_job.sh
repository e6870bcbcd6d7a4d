#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 300 python tools/sweep_time.py > gpurun_out/r02_sweep_time.log 2>&1
cat gpurun_out/r02_sweep_time.log
