#!/bin/bash
# Round-2 GPU evidence in one gpurun call (single B200):   gpurun --timeout 2400 -- 'bash tools/gpu_evidence.sh'
# Every ncu run follows a plain run of the same command that exited 0 (B200_PROFILING.md).
set -u
mkdir -p gpurun_out
O=gpurun_out
# 1. the bench lines (ours + the CPU reference arm)
python bench.py > $O/r02_bench.json 2> $O/r02_bench.err
python bench.py --impl reference > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
# 2. every launch of a short bench run with its device time (shares, not absolutes)
SHORT="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1"
$SHORT > $O/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches.csv $SHORT > $O/ncu_launches.log 2>&1
# 3. DRAM bytes of the shipped kernels per feature length (traffic vs algorithmic bytes)
TUNE="python tools/tune.py --features 32,64,128,256,512 --iters 2 --tag traffic"
$TUNE > $O/plain_traffic.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_kernel --csv --log-file $O/r02_traffic.csv $TUNE > $O/ncu_traffic.log 2>&1
# 4. full capture of the two shipped kernels at F = 128 (F = 512: r02_prof_st_f512.ncu-rep of this round)
T128="python tools/tune.py --features 128 --iters 2 --tag ncu"
$T128 > $O/plain_f128.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 2 -f -o $O/r02_prof_st_f128 $T128 > $O/ncu_f128.log 2>&1
ls -la $O/*.ncu-rep | tail -3
