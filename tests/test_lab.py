"""The experimental kernel forms (ring / fstream / fused / pull) are built only into libhgef_b200_lab.so.
Their parity tests (tests/lab/lab_forms.py) run in a subprocess that loads that library through HGEF_B200_LIB,
so the process that runs the product tests never loads it."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_product_library_refuses_experimental_forms(cuda_device):
    import torch
    from hypergef_b200 import HyperGraph, ops, synth, _native
    data = synth.make_shape("cora", seed=0)
    hg = HyperGraph(data, cuda_device, data.dataset)
    plan = ops.get_plan(hg.group_key, hg.group_row, hg.group_start, hg.group_end, hg.H_T_colind, hg.num_nodes, hg.num_edges)
    X = torch.randn(hg.num_nodes, 32, device=cuda_device)
    for flags in (_native.HG_FORCE_FUSED, _native.HG_FORCE_PULL, _native.HG_FORCE_RING, _native.HG_FORCE_FSTREAM):
        with pytest.raises(ValueError):
            ops.aggregate(plan, X, flags=flags)


def test_lab_forms_in_lab_library(cuda_device):
    lab = os.path.join(ROOT, "hypergef_b200", "libhgef_b200_lab.so")
    if not os.path.exists(lab):
        subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "hypergef_b200", "csrc"), "lab"], check=True)
    env = dict(os.environ, HGEF_B200_LIB=lab)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "lab", "lab_forms.py"), "-x", "-q",
                        "-o", "python_files=lab_*.py", "-p", "no:cacheprovider"], env=env, cwd=ROOT,
                       capture_output=True, text=True, timeout=3000)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
