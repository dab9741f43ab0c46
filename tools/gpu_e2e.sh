#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "host_pipeline" 2>&1 | tail -3
for slab in 32 64 128 256; do
  HGEF_COL_SLAB=$slab timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_slab$slab.json 2> gpurun_out/bench_slab$slab.err
  python -c "import json;d=json.loads(open('gpurun_out/bench_slab$slab.json').read().strip().splitlines()[-1]);print('slab',$slab,'value',round(d['value'],1),'e2e',d['e2e'])"
done
