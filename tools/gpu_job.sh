#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -m gpu -k "speculative or empty_batch" 2>&1 | grep -E "passed|failed|FAILED|^E " | head -20
