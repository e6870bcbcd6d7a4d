#!/bin/bash
# Collects the round's ncu evidence on one B200: launch lists per workload and full captures of the top kernels.
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
R=r02
for w in cfg2 cfg3 cfg3b cfg4 cfg4b cfg5; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_$w.csv python tools/run_workload_once.py $w 16 > /dev/null 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${R}_launches_bench_default.csv python bench.py --steps 2 --warmup 3 --min-seconds 0.01 --no-per-config --no-e2e --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
# full capture -> text summary (the .ncu-rep stays on the box: gpurun_out is capped at 64 MiB)
prof() { timeout 250 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o /tmp/${R}_$4 python tools/run_workload_once.py $1 > /dev/null 2>&1; bash tools/profile_summary.sh /tmp/${R}_$4.ncu-rep $5 gpurun_out/${R}_final_$4.txt; }
prof cfg2 biquad_cascade_x2 1 cascade 67108864
prof cfg4 dynamics_kernel 1 dynamics 67108864
prof cfg4b dynamics_spec 1 dynamics_spec 67108864
prof cfg3b fir_ols_kernel 1 ols4096 134217728
prof cfg3 reverb_ir_kernel 1 reverb_ir 98304000
prof cfg3 fir_spectrum_kernel 2 hspec4096 51072000
prof cfg3 fir_xspec_kernel 2 xspec4096 69730304
prof cfg3 fir_mac2_kernel 2 mac2_24 69730304
prof cfg3 fir_inv_kernel 2 inv4096 69730304
ls -la gpurun_out | grep ${R}_ | tail -30
