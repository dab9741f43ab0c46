#!/bin/bash
# first GPU pass for the stream form: parity tests, then timing sweeps
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "stream" > gpurun_out/st_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/st_pytest.log
tail -5 gpurun_out/st_pytest.log
timeout 300 python tools/tune.py --force-stream --tag st --sweep "FUSED=0;FUSED=1;FUSED=1,LAG=0;FUSED=1,LAG=64;FUSED=1,LAG=512;FUSED=0,ONLY=1;FUSED=0,ONLY=2" > gpurun_out/st_tune1.log 2>&1
cat gpurun_out/st_tune1.log
timeout 300 python tools/tune.py --force-stream --tag st --features 128,512 --sweep "FUSED=1,CTAS=1;FUSED=1,CTAS=2;FUSED=1,L=16;FUSED=1,L=32;FUSED=1,L=128;FUSED=1,SLAB=128;FUSED=0,CTAS=2;FUSED=0,L=128;FUSED=0,L=32" > gpurun_out/st_tune2.log 2>&1
cat gpurun_out/st_tune2.log
HGEF_NO_STREAM=1 timeout 200 python tools/tune.py --tag old > gpurun_out/st_tune_old.log 2>&1
cat gpurun_out/st_tune_old.log
