#!/bin/bash
mkdir -p gpurun_out
export HGEF_ST_RK=16 HGEF_ST_NW=14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring_kernel -s 6 -c 2 -o gpurun_out/prof_ring_f128 -f python tools/tune.py --force-stream --features 128 --iters 1 > gpurun_out/ncu_ring_f128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ring_kernel -s 6 -c 2 -o gpurun_out/prof_ring_f512 -f python tools/tune.py --force-stream --features 512 --iters 1 > gpurun_out/ncu_ring_f512.log 2>&1
tail -2 gpurun_out/ncu_ring_f512.log
