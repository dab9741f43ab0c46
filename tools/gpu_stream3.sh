#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/st_pytest_full.log 2>&1
echo "pytest rc=$?" >> gpurun_out/st_pytest_full.log
tail -3 gpurun_out/st_pytest_full.log
timeout 600 python tools/tune.py --force-stream --tag pipe --sweep "PIPE=1;PIPE=0;;PIPE=0,L=32;PIPE=1,L=32" > gpurun_out/st_tune8.log 2>&1
cat gpurun_out/st_tune8.log
timeout 300 python tools/tune.py --tag auto > gpurun_out/st_tune_auto.log 2>&1
cat gpurun_out/st_tune_auto.log
