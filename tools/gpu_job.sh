#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg5.csv python tools/run_workload_once.py cfg5 16 > /dev/null 2>&1
tail -2 gpurun_out/r02_tests_final.log; tail -2 gpurun_out/r02_bench_final.err; tail -1 gpurun_out/r02_smoke.log
