// hgef_stage.cu -- the two halves of the aggregation as separate launches, for the
// vertex/hyperedge-PARTITIONED multi-GPU path.
//
// On one GPU the hyperedge feature never leaves the SM (hgef_fused.cu).  When the vertices of a
// hyperedge live on several GPUs its feature has to be completed across ranks, so for those
// BOUNDARY hyperedges only, the two stages run as separate kernels around the exchange:
//   hg_edge_reduce   P[e,:]  = sum_{u in row e} a_in[u] * X[u,:]          (stage 1, plain stores)
//   hg_edge_scatter  Y[v,:] += a_out[v] * scale[e] * Q[e,:]  for v in row e  (stage 2, red.v4)
// over a CSR whose rows are the boundary hyperedges restricted to the local vertex block.  The same
// scatter kernel with a one-entry-per-row CSR adds received partial rows into the owner's buffer.
#include "hgef_aggr.cuh"

namespace hg {
namespace {
using namespace dev;

struct SArgs {
  const int32_t *indptr, *indices;
  const float *X, *Q, *scale, *a_out, *a_in;
  float *P, *Y;
  int64_t nrow;
  int32_t F, lpr;
};

template <int VPL>
__global__ void __launch_bounds__(kThreads) edge_reduce_kernel(const SArgs s) {
  const int lane = threadIdx.x & 31;
  const int lpr = s.lpr, groups = 32 / lpr, grp = lane / lpr;
  const int col = (lane & (lpr - 1)) * 4;
  const int F = s.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t e = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); e < s.nrow; e += nwarps) {
    const int32_t lo = __ldg(s.indptr + e), hi = __ldg(s.indptr + e + 1);
    Acc<VPL> acc;
    acc.zero();
    for (int32_t base = lo; base < hi; base += 32) {
      const int n = min(32, hi - base);
      int32_t my_v = 0;
      float my_a = 1.0f;
      if (lane < n) {
        my_v = __ldg(s.indices + base + lane);
        if (s.a_in) my_a = __ldg(s.a_in + my_v);
      }
#pragma unroll 4
      for (int r0 = 0; r0 < n; r0 += groups) {
        const int r = r0 + grp;
        const int32_t v = __shfl_sync(kFull, my_v, r & 31);
        const float w = __shfl_sync(kFull, my_a, r & 31);
        if (r < n) {
          const float *xp = s.X + (int64_t)v * F + col;
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
    for (int off = lpr; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        acc.v[j].x += __shfl_xor_sync(kFull, acc.v[j].x, off);
        acc.v[j].y += __shfl_xor_sync(kFull, acc.v[j].y, off);
        acc.v[j].z += __shfl_xor_sync(kFull, acc.v[j].z, off);
        acc.v[j].w += __shfl_xor_sync(kFull, acc.v[j].w, off);
      }
    }
    if (lane < lpr) {
      float *pp = s.P + e * F + col;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (col + j * 128 < F) *reinterpret_cast<float4 *>(pp + j * 128) = acc.v[j];
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(kThreads) edge_scatter_kernel(const SArgs s) {
  const int lane = threadIdx.x & 31;
  const int lpr = s.lpr, groups = 32 / lpr, grp = lane / lpr;
  const int col = (lane & (lpr - 1)) * 4;
  const int F = s.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t e = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); e < s.nrow; e += nwarps) {
    const int32_t lo = __ldg(s.indptr + e), hi = __ldg(s.indptr + e + 1);
    if (lo == hi) continue;
    const float sc = s.scale ? __ldg(s.scale + e) : 1.0f;
    const float *qp = s.Q + e * F + col;
    Acc<VPL> acc;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      acc.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + j * 128 < F) acc.v[j] = __ldg(reinterpret_cast<const float4 *>(qp + j * 128));
    }
    scale_acc<VPL>(acc, sc);
    for (int32_t base = lo; base < hi; base += 32) {
      const int n = min(32, hi - base);
      int32_t my_v = 0;
      float my_o = 1.0f;
      if (lane < n) {
        my_v = __ldg(s.indices + base + lane);
        if (s.a_out) my_o = __ldg(s.a_out + my_v);
      }
#pragma unroll 4
      for (int r0 = 0; r0 < n; r0 += groups) {
        const int r = r0 + grp;
        const int32_t v = __shfl_sync(kFull, my_v, r & 31);
        const float o = __shfl_sync(kFull, my_o, r & 31);
        if (r < n) {
          float *yp = s.Y + (int64_t)v * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (col + j * 128 < F)
              red_add_v4(yp + j * 128, make_float4(acc.v[j].x * o, acc.v[j].y * o, acc.v[j].z * o, acc.v[j].w * o));
        }
      }
    }
  }
}

// any-F scalar twins (F % 4 != 0 or unaligned pointers)
__global__ void __launch_bounds__(kThreads) edge_reduce_scalar_kernel(const SArgs s) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t e = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); e < s.nrow; e += nwarps) {
    const int32_t lo = s.indptr[e], hi = s.indptr[e + 1];
    for (int k = lane; k < s.F; k += 32) {
      float acc = 0.f;
      for (int32_t p = lo; p < hi; ++p) {
        const int32_t v = __ldg(s.indices + p);
        const float x = __ldg(s.X + (int64_t)v * s.F + k);
        acc = s.a_in ? fmaf(__ldg(s.a_in + v), x, acc) : acc + x;
      }
      s.P[e * s.F + k] = acc;
    }
  }
}

__global__ void __launch_bounds__(kThreads) edge_scatter_scalar_kernel(const SArgs s) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t e = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); e < s.nrow; e += nwarps) {
    const int32_t lo = s.indptr[e], hi = s.indptr[e + 1];
    const float sc = s.scale ? s.scale[e] : 1.0f;
    for (int k = lane; k < s.F; k += 32) {
      const float q = s.Q[e * s.F + k] * sc;
      for (int32_t p = lo; p < hi; ++p) {
        const int32_t v = __ldg(s.indices + p);
        atomicAdd(s.Y + (int64_t)v * s.F + k, s.a_out ? q * __ldg(s.a_out + v) : q);
      }
    }
  }
}

inline int lanes_per_row(int F) {
  int need = (F < 128 ? F : 128) / 4, l = 1;
  while (l < need) l <<= 1;
  return l;
}
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

unsigned grid_for(int64_t nrow) {
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = ceil_div<int64_t>(nrow, kWarpsPerBlock), cap = (int64_t)sm * 8;
  return (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

int hg_edge_reduce(int64_t nrow, const int32_t *d_indptr, const int32_t *d_indices, const float *d_X,
                   const float *d_a_in, float *d_P, int32_t F, int device, void *stream) {
  HG_REQUIRE(nrow >= 0 && F >= 1, "edge_reduce: bad sizes");
  if (nrow == 0) return HG_OK;
  HG_REQUIRE(d_indptr && d_indices && d_X && d_P, "edge_reduce: a required pointer is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "edge_reduce: cannot select device %d", device);
  SArgs s{};
  s.indptr = d_indptr; s.indices = d_indices; s.X = d_X; s.a_in = d_a_in; s.P = d_P;
  s.nrow = nrow; s.F = F; s.lpr = lanes_per_row(F);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(nrow);
  if (F % 4 == 0 && F <= 512 && aligned16(d_X) && aligned16(d_P)) {
    if (F <= 128) edge_reduce_kernel<1><<<grid, kThreads, 0, st>>>(s);
    else if (F <= 256) edge_reduce_kernel<2><<<grid, kThreads, 0, st>>>(s);
    else edge_reduce_kernel<4><<<grid, kThreads, 0, st>>>(s);
  } else {
    edge_reduce_scalar_kernel<<<grid, kThreads, 0, st>>>(s);
  }
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int hg_edge_scatter(int64_t nrow, const int32_t *d_indptr, const int32_t *d_indices, const float *d_Q,
                    const float *d_scale, const float *d_a_out, float *d_Y, int32_t F, int device,
                    void *stream) {
  HG_REQUIRE(nrow >= 0 && F >= 1, "edge_scatter: bad sizes");
  if (nrow == 0) return HG_OK;
  HG_REQUIRE(d_indptr && d_indices && d_Q && d_Y, "edge_scatter: a required pointer is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "edge_scatter: cannot select device %d", device);
  SArgs s{};
  s.indptr = d_indptr; s.indices = d_indices; s.Q = d_Q; s.scale = d_scale; s.a_out = d_a_out; s.Y = d_Y;
  s.nrow = nrow; s.F = F; s.lpr = lanes_per_row(F);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(nrow);
  if (F % 4 == 0 && F <= 512 && aligned16(d_Q) && aligned16(d_Y)) {
    if (F <= 128) edge_scatter_kernel<1><<<grid, kThreads, 0, st>>>(s);
    else if (F <= 256) edge_scatter_kernel<2><<<grid, kThreads, 0, st>>>(s);
    else edge_scatter_kernel<4><<<grid, kThreads, 0, st>>>(s);
  } else {
    edge_scatter_scalar_kernel<<<grid, kThreads, 0, st>>>(s);
  }
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int hg_copy_columns(void *dst, const void *src, int64_t nrow, int64_t ncol, int64_t ld_host, int64_t col0,
                    int32_t to_device, int device, void *stream) {
  HG_REQUIRE(dst != nullptr && src != nullptr, "copy_columns: NULL buffer");
  HG_REQUIRE(nrow >= 0 && ncol >= 1 && col0 >= 0 && col0 + ncol <= ld_host, "copy_columns: columns [%lld, %lld) outside a row of %lld",
             (long long)col0, (long long)(col0 + ncol), (long long)ld_host);
  if (nrow == 0) return HG_OK;
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "copy_columns: cannot select device %d", device);
  const size_t w = (size_t)ncol * sizeof(float), hp = (size_t)ld_host * sizeof(float);
  if (to_device)
    HG_CUDA_TRY(cudaMemcpy2DAsync(dst, w, static_cast<const char *>(src) + (size_t)col0 * sizeof(float), hp, w, (size_t)nrow,
                                  cudaMemcpyHostToDevice, (cudaStream_t)stream));
  else
    HG_CUDA_TRY(cudaMemcpy2DAsync(static_cast<char *>(dst) + (size_t)col0 * sizeof(float), hp, src, w, w, (size_t)nrow,
                                  cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return HG_OK;
}

}  // extern "C"
