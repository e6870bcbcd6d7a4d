#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
( time python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -6
