#!/bin/bash
# Final-tree validation on one B200: GPU tests, smoke, the default bench line, the reference arm, the cfg5 launch list.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
R=${1:-r02}
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${R}_tests_final.log
( time timeout 120 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -6 | tee gpurun_out/${R}_smoke.log
timeout 400 python bench.py > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench_final.err; echo "reference rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_cfg5.csv python tools/run_workload_once.py cfg5 16 > /dev/null 2>&1
# launch list of the bench command itself (kernels inside the CUDA-graph replays are listed by ncu)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --no-per-config --no-e2e --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/%s_bench_final.json" % "r02").read().strip().splitlines()[-1])
print("main", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"] if d.get("e2e") else None, d["gpu_launches"])
for k, v in d.get("per_config", {}).items():
    print(k, round(v["ms_per_step"], 4), round(v["roofline"]["frac"], 3))
PY
python tools/launch_agg.py gpurun_out/${R}_launches_cfg5.csv 3 2>/dev/null | head -8
python tools/launch_agg.py gpurun_out/${R}_launches_bench_default.csv 1 2>/dev/null | head -8
# memcheck over the tests of the kernels added last (packed multiply-accumulate, source-reading cascade); bounded
if [ "$2" = "memcheck" ]; then
  timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_fullsize_gpu.py tests/test_parity_gpu.py -m gpu -x -q \
    -k "source_fold or cascade_ex or parameter_rows or (long_filter_partition and (20000 or 16385 or 60000))" > gpurun_out/${R}_memcheck2.log 2>&1
  echo "memcheck rc=$?"; tail -4 gpurun_out/${R}_memcheck2.log
fi
