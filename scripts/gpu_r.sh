#!/bin/bash
# Round-2 closing check: reference arm at C5 (as rank 0 of an N=2 launch), GPU tests, smoke, default bench.
mkdir -p gpurun_out
( time WORLD_SIZE=2 RANK=0 timeout 840 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 \
    > gpurun_out/r_ref_c5.json 2> gpurun_out/r_ref_c5.err ) 2> gpurun_out/r_ref_c5.time
echo "reference c5 rc=$?"; tail -3 gpurun_out/r_ref_c5.time; head -c 1500 gpurun_out/r_ref_c5.json; echo
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
( time timeout 900 python bench.py --no-anchor > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err ) 2> gpurun_out/r_bench.time
echo "bench rc=$?"; tail -3 gpurun_out/r_bench.time; head -c 600 gpurun_out/r_bench.json; echo
