#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2_final.json 2> gpurun_out/r02_bench_n2_final.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --workload cfg2 > gpurun_out/r02_bench_cfg2_n2.json 2> gpurun_out/r02_bench_cfg2_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err
timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "broadcast" 2>&1 | tail -2
for f in r02_bench_n2_final r02_bench_cfg2_n2 r02_bench_ref_n2; do grep '^{' gpurun_out/$f.json | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$f', d.get('impl'), d['n_gpus'], d['ms_per_step'], d['value'], d['scaling'], (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('frac'))"; tail -2 gpurun_out/$f.err | cut -c1-200; done
