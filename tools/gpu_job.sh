#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
rm -f gpurun_out/r02_fft_time.log
for lib in libgrafx_b200.so libgfx_tw15.so; do
GRAFX_B200_LIB=$PWD/grafx_b200/lib/$lib timeout 300 python tools/fft_time.py >> gpurun_out/r02_fft_time.log 2>&1
done
cat gpurun_out/r02_fft_time.log
