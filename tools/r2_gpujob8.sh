#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for f in tests/test_*.py; do timeout 900 python -m pytest $f -x -q -m gpu 2>&1 | tail -1 | sed "s|^|$f: |"; done
