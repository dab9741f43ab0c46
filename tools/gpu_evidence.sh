#!/bin/bash
# Round evidence on one B200 (run through gpurun): tests, bench (both arms), launch list, ncu captures, per-F traffic.
# Everything lands in gpurun_out/; tools/collect_evidence.py copies the summaries into profiles/.
mkdir -p gpurun_out
T=${1:-r01b}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -2 gpurun_out/${T}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python -c "import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print('value',round(d['value'],1),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'],d['clocks'])"
# launch list of the same command (short), then one full capture per stage kernel at F=512 and F=128
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/${T}_ncu_launches.log 2>&1
for F in 512 128 32; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 6 -c 2 -o gpurun_out/${T}_prof_stream_f$F -f \
      python tools/tune.py --features $F --iters 1 > gpurun_out/${T}_ncu_f$F.log 2>&1
done
# DRAM bytes of one call per F (all stream kernels of the call)
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:stream_kernel --csv \
    --log-file gpurun_out/${T}_traffic.csv python tools/tune.py --iters 1 > gpurun_out/${T}_ncu_traffic.log 2>&1
# other configurations with the shipped defaults
timeout 600 python tools/tune.py --tag default > gpurun_out/${T}_tune_default.log 2>&1
timeout 600 python tools/tune.py --shape dblp --replicas 30 --features 128 --tag dblp >> gpurun_out/${T}_tune_default.log 2>&1
timeout 600 python tools/tune.py --shape walmart --replicas 8 --features 32,128 --tag walmart >> gpurun_out/${T}_tune_default.log 2>&1
HGEF_NO_STREAM=1 timeout 600 python tools/tune.py --shape walmart --replicas 8 --features 32,128 --tag walmart_nostream >> gpurun_out/${T}_tune_default.log 2>&1
cat gpurun_out/${T}_tune_default.log
timeout 600 python tools/fwd_bwd_bench.py > gpurun_out/${T}_fwd_bwd.json 2>&1; tail -1 gpurun_out/${T}_fwd_bwd.json
timeout 600 python tools/epoch_bench.py > gpurun_out/${T}_epoch.json 2>&1; tail -1 gpurun_out/${T}_epoch.json
timeout 900 python tools/c5_single.py > gpurun_out/${T}_c5_single.log 2>&1; tail -1 gpurun_out/${T}_c5_single.log
