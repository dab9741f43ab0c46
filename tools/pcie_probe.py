#!/usr/bin/env python
"""PCIe staging rates on this box: contiguous vs column-slab (cudaMemcpy2DAsync) copies, one direction and both."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hypergef_b200 import _native

N, F = 1261888, 512
dev = torch.device("cuda:0")
hx = torch.empty(N, F).pin_memory(); hy = torch.empty(N, F).pin_memory()
hx.normal_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

gb = N * F * 4 / 1e9
dx = torch.empty(N, F, device=dev); dy = torch.empty(N, F, device=dev)
def up_whole():
    with torch.cuda.stream(s1): dx.copy_(hx, non_blocking=True)
def down_whole():
    with torch.cuda.stream(s2): hy.copy_(dy, non_blocking=True)
print(f"contiguous H2D {gb / timed(up_whole) * 1e3:6.1f} GB/s   D2H {gb / timed(down_whole) * 1e3:6.1f} GB/s   both "
      f"{2 * gb / timed(lambda: (up_whole(), down_whole())) * 1e3:6.1f} GB/s (sum)")
for w in (32, 64, 128, 256):
    d = torch.empty(N, w, device=dev)
    def up():
        for c0 in range(0, F, w):
            _native.call("hg_copy_columns", d.data_ptr(), hx.data_ptr(), N, w, F, c0, 1, 0, s1.cuda_stream)
    def down():
        for c0 in range(0, F, w):
            _native.call("hg_copy_columns", hy.data_ptr(), d.data_ptr(), N, w, F, c0, 0, 0, s2.cuda_stream)
    print(f"slab {w:4d} cols   H2D {gb / timed(up) * 1e3:6.1f} GB/s   D2H {gb / timed(down) * 1e3:6.1f} GB/s   both "
          f"{2 * gb / timed(lambda: (up(), down())) * 1e3:6.1f} GB/s (sum)")
