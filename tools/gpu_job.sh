#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -x -k "ballistics or cfg4" 2>&1 | tail -5 > gpurun_out/r02_ball.log
timeout 300 python tools/ball_time.py >> gpurun_out/r02_ball.log 2>&1
cat gpurun_out/r02_ball.log
