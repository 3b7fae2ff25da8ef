#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python scripts/bench_potrf_variants.py 1024 4096 8424 16384 2>&1 | tail -1 | tee gpurun_out/j_variants.jsonl
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/j_pytest.txt; tail -3 gpurun_out/j_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -3
timeout 300 python scripts/run_configs.py c2 c4 2>&1 | tail -2 | tee gpurun_out/j_configs.jsonl
