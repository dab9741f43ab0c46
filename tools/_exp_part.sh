W=${1:-2}
for extra in "" "--single-stage-a"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29533 tools/run_partition.py --scale 0.2 --F 256 --iters 10 --check 0 $extra 2>&1 | grep "^{" | sed "s/^/[$W gpus $extra] /" | tee -a gpurun_out/part_${W}gpu.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29534 tools/run_partition.py --scale 0.01 --F 64 --iters 2 --check 1 --single-stage-a 2>&1 | grep "^{" | sed "s/^/[$W gpus check single-stage-a] /" | tee -a gpurun_out/part_${W}gpu.log
