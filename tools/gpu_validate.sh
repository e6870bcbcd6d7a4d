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
python - <<'PY'
import json
d = json.loads(open("gpurun_out/%s_bench_final.json" % "r02").read().strip().splitlines()[-1])
print("main", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"] if d.get("e2e") else None, d["gpu_launches"])
for k, v in d.get("per_config", {}).items():
    print(k, round(v["ms_per_step"], 4), round(v["roofline"]["frac"], 3))
PY
python tools/launch_agg.py gpurun_out/${R}_launches_cfg5.csv 3 | head -14
