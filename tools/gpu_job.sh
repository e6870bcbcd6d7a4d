#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
rm -f gpurun_out/r02_conv_time.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -x -k "fir or reverb or conv" 2>&1 | tail -2 > gpurun_out/r02_t2.log
for lib in libgrafx_b200.so libgfx_peel4.so libgfx_peel12.so; do
GRAFX_B200_LIB=$PWD/grafx_b200/lib/$lib timeout 300 python tools/conv_time.py 2>&1 | head -1 >> gpurun_out/r02_conv_time.log
done
cat gpurun_out/r02_t2.log gpurun_out/r02_conv_time.log
