export HGEF_B200_LIB=hypergef_b200/libhgef_b200_lab.so
P="fs_split=2,fs_pol_x=0"
timeout 600 python tools/tune.py --force-fstream --verify --tag alt5 --features 256 --sweep "$P,fs_item_kb=32;$P,fs_item_kb=16;$P,fs_item_kb=12;$P,fs_item_kb=8;$P,fs_item_kb=4;$P,fs_item_kb=8,fs_debug=1;$P,fs_item_kb=8,fs_debug=3;$P,fs_item_kb=16,fs_debug=3;$P,fs_item_kb=8,fs_occ=2;$P,fs_item_kb=8,fs_pipe=0;$P,fs_item_kb=8,fs_discard=0" > gpurun_out/alt5_tune.log 2>&1; echo tune rc=$?
cat gpurun_out/alt5_tune.log | cut -c1-300
timeout 300 python tools/tune.py --force-fstream --verify --tag alt5 --features 128,512 --sweep "$P,fs_item_kb=8,fs_slab=256;$P,fs_item_kb=16,fs_slab=256;$P,fs_item_kb=16;$P,fs_item_kb=8" 2>&1 | cut -c1-300 | tee gpurun_out/alt5_tune2.log
T="python tools/tune.py --force-fstream --tag ncu --features 256 --iters 1 --sweep $P,fs_item_kb=8;$P,fs_item_kb=12;$P,fs_item_kb=16"
$T > gpurun_out/alt5_plain.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:astream_kernel --csv --log-file gpurun_out/alt5_traffic.csv $T > gpurun_out/alt5_ncu.log 2>&1
grep -c astream gpurun_out/alt5_traffic.csv
