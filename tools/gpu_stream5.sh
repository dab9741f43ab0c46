#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "stream_form or kernel_form" > gpurun_out/st_pytest3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/st_pytest3.log
tail -3 gpurun_out/st_pytest3.log
timeout 600 python tools/tune.py --tag static --sweep "STATIC=0;;OCC=4;L=16;L=16,OCC=4;L=48;L=64;ONLY=1;ONLY=2;OCC=4,ONLY=1;OCC=4,ONLY=2" > gpurun_out/st_tune13.log 2>&1
cat gpurun_out/st_tune13.log
