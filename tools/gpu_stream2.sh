#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "stream_form" > gpurun_out/st_pytest2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/st_pytest2.log
tail -3 gpurun_out/st_pytest2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 2 -o gpurun_out/prof_st2_f128 -f python tools/tune.py --force-stream --features 128 --iters 1 > gpurun_out/ncu_st2_f128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 2 -o gpurun_out/prof_st2_f32 -f python tools/tune.py --force-stream --features 32 --iters 1 > gpurun_out/ncu_st2_f32.log 2>&1
tail -2 gpurun_out/ncu_st2_f32.log
