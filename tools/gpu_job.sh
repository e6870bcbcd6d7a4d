#!/bin/bash
# scratch driver for one gpurun call (edited per call); outputs under gpurun_out/
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/r02_t1.log
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_fullsize_gpu.py 2>&1 | tail -25 > gpurun_out/r02_tests_all.log
timeout 600 python bench.py > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err
tail -5 gpurun_out/r02_t1.log; tail -5 gpurun_out/r02_tests_all.log; tail -c 1500 gpurun_out/r02_bench1.err; head -c 3000 gpurun_out/r02_bench1.json
