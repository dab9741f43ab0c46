#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/tune.py --force-stream --tag kv --sweep ";OCC=2;OCC=2,PIPE=1;OCC=2,L=128;OCC=2,ONLY=1;OCC=2,ONLY=2;OCC=2,SW=16;OCC=2,PIPE=1,SW=16" > gpurun_out/st_tune11.log 2>&1
cat gpurun_out/st_tune11.log
