"""``balance_schedule`` -- drop-in for ``HyperGsys/balancer.py:4-33``.

Same constructor, same four attributes (``balan_key``, ``balan_row``, ``group_st``,
``group_ed``) with bit-identical contents, computed by the native balancer
(``hg_balance_*``, hypergef_b200/csrc/hgef_balancer.cu) instead of an O(G) interpreted loop.

* CPU ``H_T_csrptr`` (what the reference passes, ``hypergraph.py:77``) -> host balancer,
  attributes are int32 numpy arrays (they index, slice, ``len()`` and convert with
  ``torch.Tensor(...)`` like the reference's lists).
* CUDA ``H_T_csrptr`` -> device balancer, attributes are int32 CUDA tensors (the native
  builder the reference left as a TODO, ``source/balancer/balancer_kernel.cu:34``).

An all-empty matrix raises ``IndexError`` exactly like ``balancer.py:32``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native

__all__ = ["balance_schedule"]


class balance_schedule:  # noqa: N801  (reference spelling)
    def __init__(self, ngs, H_T_csrptr):
        if not isinstance(H_T_csrptr, torch.Tensor):
            H_T_csrptr = torch.as_tensor(np.asarray(H_T_csrptr))
        if H_T_csrptr.dim() != 1 or H_T_csrptr.numel() < 1:
            raise ValueError("H_T_csrptr must be a 1-D offset array with at least one entry")
        if H_T_csrptr.dtype not in (torch.int32, torch.int64):
            raise TypeError(f"H_T_csrptr must be int32/int64, got {H_T_csrptr.dtype}")
        ngs = int(ngs)
        self.ngs = ngs
        self.nrow = H_T_csrptr.shape[0] - 1
        if H_T_csrptr.dtype == torch.int64:
            if self.nrow >= 0 and int(H_T_csrptr[-1]) > np.iinfo(np.int32).max:
                raise ValueError("H_T_csrptr exceeds the int32 range of the index arrays")
            H_T_csrptr = H_T_csrptr.to(torch.int32)
        ptr = H_T_csrptr.contiguous()
        nkey, ngroup = C.c_int64(), C.c_int64()
        if ptr.is_cuda:
            dev = ptr.device.index if ptr.device.index is not None else torch.cuda.current_device()
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.call("hg_balance_count_dev", self.nrow, ptr.data_ptr(), ngs, C.byref(nkey),
                         C.byref(ngroup), dev, stream)
            key = torch.empty(nkey.value, dtype=torch.int32, device=ptr.device)
            row, st, ed = (torch.empty(ngroup.value, dtype=torch.int32, device=ptr.device) for _ in range(3))
            _native.call("hg_balance_fill_dev", self.nrow, ptr.data_ptr(), ngs, nkey.value, ngroup.value,
                         key.data_ptr(), row.data_ptr(), st.data_ptr(), ed.data_ptr(), dev, stream)
        else:
            pn = ptr.numpy()
            _native.call("hg_balance_count_host", self.nrow, pn.ctypes.data, ngs, C.byref(nkey),
                         C.byref(ngroup))
            key = np.empty(nkey.value, np.int32)
            row, st, ed = (np.empty(ngroup.value, np.int32) for _ in range(3))
            _native.call("hg_balance_fill_host", self.nrow, pn.ctypes.data, ngs, key.ctypes.data,
                         row.ctypes.data, st.ctypes.data, ed.ctypes.data)
        self.balan_key, self.balan_row, self.group_st, self.group_ed = key, row, st, ed
        self.work_p_sum = int(nkey.value - 1)          # balancer.py:31 running total = #segments
