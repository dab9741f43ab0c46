timeout 900 python -m pytest tests/test_convs.py -x -q -m gpu -k "stage_entry or projected_layer or two_layer" > gpurun_out/n1_pytest.log 2>&1; echo pytest rc=$?; tail -5 gpurun_out/n1_pytest.log | cut -c1-400
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/n1_layer.log
import json, torch, sys
sys.path.insert(0, '.')
import bench
import hypergef_b200 as hgef
from hypergef_b200 import synth
dev = torch.device('cuda:0')
data = synth.make_shape('pubmed', replicas=64, seed=0, device=dev)
hg = hgef.HyperGraph(data, dev, data.dataset)
W = torch.ones(hg.num_edges, device=dev)
print(json.dumps(bench.layer_block(hg, W)))
PY
