#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "dynamics or cfg4 or envelope or approx or compressor or noisegate or empty_batch or slow_pole" 2>&1 | tail -2
