#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "training_mode or render or backward or captured" 2>&1 | tail -12
