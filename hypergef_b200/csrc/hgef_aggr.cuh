// hgef_aggr.cuh -- device building blocks shared by the aggregation kernels.
#pragma once

#include "hgef_plan.cuh"

namespace hg {
namespace dev {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr unsigned kFull = 0xffffffffu;

struct Args {
  const int32_t *key, *colind, *seg_edge, *seg_slot, *seg_list;  // seg_list: optional indirection
  const int32_t *row, *st, *ed;                                   // group schedule only
  const float *X, *s1, *s2, *a_out, *a_in;
  float *Y, *scratch;
  int64_t nwork;   // segments (or listed segments, or groups)
  int32_t F;
  int32_t lpr;     // lanes per feature row (power of two <= 32); 32/lpr rows per warp step
};

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float edge_scale(const Args &a, int32_t e) {
  float s = 1.0f;
  if (a.s1) s = __ldg(a.s1 + e);
  if (a.s2) s *= __ldg(a.s2 + e);
  return s;
}

// ---- 128-bit path: F % 4 == 0.  Lane l holds columns (l % lpr)*4 + j*128 .. +3, j < VPL ----
template <int VPL>
struct Acc {
  float4 v[VPL];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int j = 0; j < VPL; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
};

// sum_{p in [lo,hi)} a_in[v_p] * X[v_p, cols]; every lane ends with the full sum of its columns
template <int VPL>
__device__ __forceinline__ void gather_rows(const Args &a, int32_t lo, int32_t hi, int lane, int col,
                                            Acc<VPL> &acc) {
  const int groups = 32 / a.lpr;
  const int grp = lane / a.lpr;
  const int F = a.F;
  for (int32_t base = lo; base < hi; base += 32) {
    const int n = min(32, hi - base);
    int32_t my_v = 0;
    float my_a = 1.0f;
    if (lane < n) {
      my_v = __ldg(a.colind + base + lane);
      if (a.a_in) my_a = __ldg(a.a_in + my_v);
    }
#pragma unroll 4
    for (int r0 = 0; r0 < n; r0 += groups) {
      const int r = r0 + grp;
      const int32_t v = __shfl_sync(kFull, my_v, r & 31);
      const float w = __shfl_sync(kFull, my_a, r & 31);
      if (r < n) {
        const float *xp = a.X + (int64_t)v * F + col;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          if (col + j * 128 < F) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(xp + j * 128));
            acc.v[j].x = fmaf(w, x.x, acc.v[j].x);
            acc.v[j].y = fmaf(w, x.y, acc.v[j].y);
            acc.v[j].z = fmaf(w, x.z, acc.v[j].z);
            acc.v[j].w = fmaf(w, x.w, acc.v[j].w);
          }
        }
      }
    }
  }
  // butterfly across the row groups (only when a row takes fewer than 32 lanes => VPL == 1)
  for (int off = a.lpr; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      acc.v[j].x += __shfl_xor_sync(kFull, acc.v[j].x, off);
      acc.v[j].y += __shfl_xor_sync(kFull, acc.v[j].y, off);
      acc.v[j].z += __shfl_xor_sync(kFull, acc.v[j].z, off);
      acc.v[j].w += __shfl_xor_sync(kFull, acc.v[j].w, off);
    }
  }
}

// Y[v_p, cols] += acc * a_out[v_p] for p in [lo,hi)
template <int VPL>
__device__ __forceinline__ void scatter_rows(const Args &a, int32_t lo, int32_t hi, int lane, int col,
                                             const Acc<VPL> &acc) {
  const int groups = 32 / a.lpr;
  const int grp = lane / a.lpr;
  const int F = a.F;
  for (int32_t base = lo; base < hi; base += 32) {
    const int n = min(32, hi - base);
    int32_t my_v = 0;
    float my_o = 1.0f;
    if (lane < n) {
      my_v = __ldg(a.colind + base + lane);
      if (a.a_out) my_o = __ldg(a.a_out + my_v);
    }
#pragma unroll 4
    for (int r0 = 0; r0 < n; r0 += groups) {
      const int r = r0 + grp;
      const int32_t v = __shfl_sync(kFull, my_v, r & 31);
      const float o = __shfl_sync(kFull, my_o, r & 31);
      if (r < n) {
        float *yp = a.Y + (int64_t)v * F + col;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
          if (col + j * 128 < F)
            red_add_v4(yp + j * 128, make_float4(acc.v[j].x * o, acc.v[j].y * o, acc.v[j].z * o,
                                                 acc.v[j].w * o));
        }
      }
    }
  }
}

template <int VPL>
__device__ __forceinline__ void scale_acc(Acc<VPL> &acc, float s) {
#pragma unroll
  for (int j = 0; j < VPL; ++j) {
    acc.v[j].x *= s; acc.v[j].y *= s; acc.v[j].z *= s; acc.v[j].w *= s;
  }
}


}  // namespace dev
}  // namespace hg
