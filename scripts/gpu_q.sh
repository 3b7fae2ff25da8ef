#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "potrf" 2>&1 | tail -2
timeout 120 python scripts/bench_potrf_variants.py 1024 2048 4096 8424 16384 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:round(v['ms_mean'],3) for k,v in d.items() if k!='lib'})"
timeout 200 python scripts/_t_c2.py 2>&1 | tail -5
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
