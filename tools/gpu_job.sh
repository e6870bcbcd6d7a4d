#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2>/dev/null
tail -2 gpurun_out/r02_tests_final.log; tail -2 gpurun_out/r02_bench_final.err; tail -1 gpurun_out/r02_smoke.log
