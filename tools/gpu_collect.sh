#!/bin/bash
# Collects the round's evidence on one B200: GPU tests, bench lines, ncu launch lists and full captures.
mkdir -p gpurun_out; cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
for w in cfg3 cfg3b cfg4 cfg4b cfg5; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
for w in cfg3 cfg3b cfg4 cfg5; do
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$w.csv python tools/run_workload_once.py $w 16 > /dev/null 2>&1
done
prof() { timeout 250 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/final_$4 python tools/run_workload_once.py $1 > /dev/null 2>&1; }
prof cfg2 biquad_cascade 1 cascade
prof cfg4 dynamics_kernel 1 dynamics
prof cfg3b fir_ols_kernel 1 ols4096
prof cfg3 reverb_ir_kernel 1 reverb_ir
prof cfg3 fir_spectrum_kernel 2 hspec
prof cfg3 fir_xspec_kernel 2 xspec
prof cfg3 fir_mac_kernel 2 mac
prof cfg3 fir_inv_kernel 2 inv
timeout 120 python tools/backward_time.py 256 lfilter > gpurun_out/backward_time.log 2>&1
timeout 120 python tools/backward_time.py 256 fsm >> gpurun_out/backward_time.log 2>&1
timeout 120 python tools/captured_bench.py 8 1 32768 > gpurun_out/captured.log 2>&1
timeout 120 python tools/captured_bench.py 32 1 131072 >> gpurun_out/captured.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_step.csv python tools/backward_time.py 256 lfilter > /dev/null 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
ls -la gpurun_out | tail -40
