#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/e_pytest.txt; tail -30 gpurun_out/e_pytest.txt
