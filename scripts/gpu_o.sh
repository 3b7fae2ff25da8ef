#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
L=$PWD/gpar_b200
timeout 180 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "potrf" 2>&1 | tail -2
for rep in 1 2; do
for v in "" _old; do
  GPAR_B200_LIB=$L/libgpar_b200$v.so timeout 120 python scripts/bench_potrf_variants.py 1024 2048 4096 8424 16384 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v'.ljust(6), {k:round(v['ms_mean'],3) for k,v in d.items() if k!='lib'})"
done; done
