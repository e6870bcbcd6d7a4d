#!/bin/bash
# ncu evidence of the final kernels on one B200: DRAM traffic per kernel (cfg5 chunk of 32 renders, cfg3), full captures of
# the packed multiply-accumulate kernel and of the source-reading cascade, launch list of cfg3.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
R=${1:-r02}
M=dram__bytes_read.sum,dram__bytes_write.sum
timeout 200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${R}_traffic_cfg5.csv python tools/run_workload_once.py cfg5 32 > /dev/null 2>&1
timeout 150 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${R}_traffic_cfg3.csv python tools/run_workload_once.py cfg3 > /dev/null 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_cfg3.csv python tools/run_workload_once.py cfg3 > /dev/null 2>&1
prof() { timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/${R}_$4 python tools/run_workload_once.py $1 $6 > /dev/null 2>&1; bash tools/profile_summary.sh /tmp/${R}_$4.ncu-rep $5 gpurun_out/${R}_final_$4.txt; }
prof cfg3 fir_mac3_kernel 1 mac3_24 134217728
prof cfg5 biquad_cascade_x2 1 cascade_src 134217728 16
python tools/traffic_agg.py gpurun_out/${R}_traffic_cfg5.csv 3 4 | tail -3
python tools/traffic_agg.py gpurun_out/${R}_traffic_cfg3.csv 3 1 | tail -3
head -12 gpurun_out/${R}_final_mac3_24.txt
