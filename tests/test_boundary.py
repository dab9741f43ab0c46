"""The drop-in boundary: the C-ABI library loads, exports exactly what include/hgef_b200.h
declares, and the product never touches the oracle or a CPU fallback."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT
from hypergef_b200 import _native


def _declared():
    text = open(os.path.join(ROOT, "include", "hgef_b200.h")).read()
    return sorted(set(re.findall(r"HG_API\s+[\w\s\*]+?\b(hg_\w+)\s*\(", text)))


def test_header_and_binding_list_the_same_functions():
    declared = _declared()
    assert len(declared) >= 19
    assert sorted(set(_native.SIGNATURES) | {"hg_last_error"}) == declared


def test_library_exports_every_declared_symbol():
    lib = _native.lib()
    for name in _declared():
        assert getattr(lib, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert [s for s in exported if s.startswith("hg_")] == _declared()
    assert not [s for s in exported if not s.startswith("hg_")], "only the C-ABI is visible"


def test_header_is_plain_c_and_a_c_host_links(tmp_path):
    """include/hgef_b200.h compiles as C (no C++, no torch, no CUDA headers) and a C host program links against the
    library and runs the host-side balancer through it -- the stub INTEGRATION.md section 3 describes."""
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include "hgef_b200.h"
int main(void) {
  const int32_t ptr[] = {0, 5, 5, 7, 14};          /* SURVEY A4 golden vector, ngs = 3 */
  int64_t S = 0, G = 0;
  if (hg_abi_version() != 2) return 2;
  if (hg_balance_count_host(4, ptr, 3, &S, &G) != HG_OK) { puts(hg_last_error()); return 3; }
  int32_t key[16], row[16], st[16], ed[16];
  if (hg_balance_fill_host(4, ptr, 3, key, row, st, ed) != HG_OK) { puts(hg_last_error()); return 4; }
  printf("%lld %lld %d %d %d\n", (long long)S, (long long)G, key[1], key[(int)S - 1], row[4]);
  return 0;
}
''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host"
    libdir = os.path.dirname(_native.LIB_PATH)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-l:" + os.path.basename(_native.LIB_PATH), "-Wl,-rpath," + libdir],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split() == ["7", "14", "3", "14", "2"]          # nkey = S + 1, G, key[1], sentinel, balan_row[4]


def test_abi_version_and_error_channel():
    lib = _native.lib()
    assert lib.hg_abi_version() == 2
    n = ctypes.c_int64()
    rc = lib.hg_balance_count_host(1, None, 3, ctypes.byref(n), ctypes.byref(n))
    assert rc == _native.HG_EINVAL and b"csrptr" in lib.hg_last_error()
    with pytest.raises(ValueError):
        _native.check(rc)


def test_product_never_imports_the_oracle_or_the_reference():
    pkg = os.path.join(ROOT, "hypergef_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "hg_oracle" not in text and "libhgref" not in text, f
                assert "/root/reference" not in text, f


def test_ops_refuse_cpu_tensors():
    """No CPU fallback: CPU tensors are an error, not a slow path."""
    import torch
    from hypergef_b200 import ops
    i = torch.zeros(2, dtype=torch.int32)
    x = torch.zeros(2, 4)
    with pytest.raises(ValueError, match="CUDA"):
        ops.unignnaggr(i, i, i, i, i, i, x)


def test_compat_registers_reference_module_names():
    import importlib
    import sys
    from hypergef_b200 import compat
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in ("HyperGsys", "hgnnaggr", "unignnaggr")}
    try:
        compat.install(force=True)
        from HyperGsys.balancer import balance_schedule  # noqa: F401
        from HyperGsys.hypergraph import HyperGraph  # noqa: F401
        from HyperGsys.source.python.hgnnaggr import HGNNAggr  # noqa: F401
        from HyperGsys.source.python.unignnconv import UniGNNConv, UniGNNConvdeg  # noqa: F401
        ext = importlib.import_module("unignnaggr")
        assert ext.unignnconv is ext.unignnaggr and ext.unignnconvdeg is ext.unignnaggrdeg
        assert hasattr(importlib.import_module("hgnnaggr"), "hgnnaggr_max")
    finally:
        for k in [k for k in sys.modules if k.split(".")[0] in ("HyperGsys", "hgnnaggr", "unignnaggr")]:
            del sys.modules[k]
        sys.modules.update({k: v for k, v in saved.items() if v is not None})
