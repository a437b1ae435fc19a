#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/bench_reference_pattern.py 2>&1 | tail -3
