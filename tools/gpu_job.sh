#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "training or trains or without_backward or backward or next_" 2>&1 | grep -E "Error|error|assert|^E |passed|failed|FAILED" | head -40
