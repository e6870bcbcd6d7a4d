#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_tests_final.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
( time timeout 600 python bench.py ) > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_cfg5.csv python tools/run_workload_once.py cfg5 16 > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py -q -m gpu -x -k "envelope or speculative or partition_sizes or firfilter_fused or pipelined or render_one_by_one or common_parameters or fir_conv_filter_repeat" > gpurun_out/r02_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_memcheck.log
tail -3 gpurun_out/r02_tests_final.log; tail -4 gpurun_out/r02_bench_final.err; tail -2 gpurun_out/r02_smoke.log; tail -5 gpurun_out/r02_memcheck.log; head -c 600 gpurun_out/r02_bench_reference.json
