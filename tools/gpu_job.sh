#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
for lib in libgrafx_b200.so libgfx_sc.so; do echo "== $lib"; GRAFX_B200_LIB=$PWD/grafx_b200/lib/$lib timeout 300 python tools/quick_time.py 2>&1 | head -3; done
GRAFX_B200_LIB=$PWD/grafx_b200/lib/libgfx_sc.so timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "iir or cascade or cfg2 or kat or low_frequency or geq" 2>&1 | tail -2
