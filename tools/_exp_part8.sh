W=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29533 tools/run_partition.py --scale 0.2 --F 256 --iters 10 --check 0 2>&1 | grep "^{" | sed "s/^/[$W gpus scale 0.2] /" | tee -a gpurun_out/part_${W}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29534 tools/run_partition.py --scale 1.0 --F 256 --iters 5 --check 0 2>&1 | grep "^{" | sed "s/^/[$W gpus full C5] /" | tee -a gpurun_out/part_${W}gpu.log
