#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_tests_final.log
timeout 600 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:reverb_ir_kernel -s 1 -c 1 -f -o /tmp/r02_ir python tools/run_workload_once.py cfg3 > /dev/null 2>&1; bash tools/profile_summary.sh /tmp/r02_ir.ncu-rep 98304000 gpurun_out/r02_final_reverb_ir.txt
tail -2 gpurun_out/r02_tests_final.log; tail -2 gpurun_out/r02_bench_final.err; tail -1 gpurun_out/r02_smoke.log
