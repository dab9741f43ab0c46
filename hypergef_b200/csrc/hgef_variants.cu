// hgef_variants.cu -- first-stage mean / max variants and the hyperedge-weight gradient.
//
// hg_plan_max_forward / _backward: the BALANCED max (one warp per balancer segment; the segments of a split
// hyperedge meet in a packed (value, vertex) atomicMax; the second stage is the stream form's stage B, so Y is
// written once without atomics).  The mean goes through the sum kernels with degE / |e| as the hyperedge scale
// (hypergef_b200/ops.py).  What follows first are the reference's own schemes, kept as the fallback and for A/B:
// they run over the UN-balanced CSR of H^T (one warp per hyperedge, lanes across feature
// columns so every row access is a coalesced 128-byte line), as the reference's
// hgnnaggr_mean / hgnnaggr_max do (hgnnaggr_cuda.cu:86-208).  Differences on purpose:
//   * the hyperedge loop is bounded by the number of hyperedges; the reference passes
//     degV.size(0) = N (hgnnaggr_cuda.cu:419,485), which is only right when N == M;
//   * any F is accepted (the reference launches F/32 column blocks and F%32 tails are lost).
#include <cfloat>

#include "hgef_stream.cuh"

namespace hg {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

struct VArgs {
  const int32_t *indptr, *indices;
  const float *X, *G, *s1, *s2, *a_out, *a_in;
  const int32_t *record_in;
  int32_t *record_out;
  float *Y, *dW;
  int64_t nedge;
  int32_t F;
};

__device__ __forceinline__ float escale(const VArgs &a, int64_t e) {
  float s = 1.0f;
  if (a.s1) s = __ldg(a.s1 + e);
  if (a.s2) s *= __ldg(a.s2 + e);
  return s;
}

enum { kMean = 0, kMaxFwd = 1, kMaxBwd = 2 };

template <int MODE>
__global__ void __launch_bounds__(kThreads) edge_kernel(const VArgs a) {
  const int lane = threadIdx.x & 31;
  const int F = a.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t e = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); e < a.nedge; e += nwarps) {
    const int32_t lo = a.indptr[e], hi = a.indptr[e + 1];
    const float sc = escale(a, e);
    for (int k = lane; k < F; k += 32) {
      if (MODE == kMean) {
        float acc = 0.f;
        for (int32_t p = lo; p < hi; ++p) acc += __ldg(a.X + (int64_t)__ldg(a.indices + p) * F + k);
        if (hi > lo) acc *= sc / (float)(hi - lo);   // hgnnaggr_cuda.cu:107
        for (int32_t p = lo; p < hi; ++p) {
          const int32_t v = __ldg(a.indices + p);
          atomicAdd(a.Y + (int64_t)v * F + k, a.a_out ? acc * __ldg(a.a_out + v) : acc);
        }
      } else if (MODE == kMaxFwd) {
        float acc = -1e5f;                            // hgnnaggr_cuda.cu:155
        int32_t rec = 0;
        for (int32_t p = lo; p < hi; ++p) {
          const int32_t v = __ldg(a.indices + p);
          const float x = __ldg(a.X + (int64_t)v * F + k);
          if (x > acc) { acc = x; rec = v; }
        }
        acc *= sc;
        a.record_out[e * F + k] = rec;
        for (int32_t p = lo; p < hi; ++p) {
          const int32_t v = __ldg(a.indices + p);
          atomicAdd(a.Y + (int64_t)v * F + k, a.a_out ? acc * __ldg(a.a_out + v) : acc);
        }
      } else {
        float acc = 0.f;
        for (int32_t p = lo; p < hi; ++p) acc += __ldg(a.G + (int64_t)__ldg(a.indices + p) * F + k);
        acc *= sc;
        const int32_t v = a.record_in[e * F + k];
        atomicAdd(a.Y + (int64_t)v * F + k, a.a_out ? acc * __ldg(a.a_out + v) : acc);
      }
    }
  }
}

// dW[e] = s1[e] * sum_k (sum_u a_in[u] X[u,k]) * (sum_v a_out[v] G[v,k])
__global__ void __launch_bounds__(kThreads) weight_grad_kernel(const VArgs a) {
  const int lane = threadIdx.x & 31;
  const int F = a.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t e = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); e < a.nedge; e += nwarps) {
    const int32_t lo = a.indptr[e], hi = a.indptr[e + 1];
    float tot = 0.f;
    for (int k = lane; k < F; k += 32) {
      float r1 = 0.f, r2 = 0.f;
      for (int32_t p = lo; p < hi; ++p) {
        const int32_t v = __ldg(a.indices + p);
        const float x = __ldg(a.X + (int64_t)v * F + k), g = __ldg(a.G + (int64_t)v * F + k);
        r1 = a.a_in ? fmaf(__ldg(a.a_in + v), x, r1) : r1 + x;
        r2 = a.a_out ? fmaf(__ldg(a.a_out + v), g, r2) : r2 + g;
      }
      tot = fmaf(r1, r2, tot);
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
    if (lane == 0) a.dW[e] = a.s1 ? tot * __ldg(a.s1 + e) : tot;
  }
}

// ---- balanced max -----------------------------------------------------------------------------------
// order-preserving packing: larger value wins; among equal values the SMALLER vertex id wins, which is the
// reference's "first strictly greater member in ascending order" (hgnnaggr_cuda.cu:155-163)
__device__ __forceinline__ unsigned long long pack_max(float x, int32_t v) {
  uint32_t u = __float_as_uint(x);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)v);
}
__device__ __forceinline__ void unpack_max(unsigned long long p, float &x, int32_t &v) {
  uint32_t u = (uint32_t)(p >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  x = __uint_as_float(u);
  v = (int32_t)(0xffffffffu - (uint32_t)(p & 0xffffffffu));
}

struct MArgs {
  const int32_t *key, *colind, *seg_edge, *seg_slot, *heavy_segs;
  const float *X, *s1, *s2, *a_out;
  float *xe, *dX;
  unsigned long long *packed;      // [nheavy_edges, F]
  int32_t *record;
  const int32_t *record_in;
  int64_t nseg, nheavy_segs, nedge;
  int32_t F;
};

__global__ void max_init_kernel(int64_t n, unsigned long long *packed) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) packed[i] = pack_max(-1e5f, 0);   // the reference's sentinel (hgnnaggr_cuda.cu:155)
}

// one warp per balancer segment: running (max, argmax) per column over the segment's members
__global__ void __launch_bounds__(kThreads) seg_max_kernel(const MArgs a) {
  const int lane = threadIdx.x & 31;
  const int F = a.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t sgm = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); sgm < a.nseg; sgm += nwarps) {
    const int32_t lo = a.key[sgm], hi = a.key[sgm + 1], e = a.seg_edge[sgm], slot = a.seg_slot[sgm];
    float sc = 1.0f;
    if (a.s1) sc = __ldg(a.s1 + e);
    if (a.s2) sc *= __ldg(a.s2 + e);
    for (int k = lane; k < F; k += 32) {
      float best = -1e5f;
      int32_t rec = 0;
      for (int32_t p = lo; p < hi; ++p) {
        const int32_t v = __ldg(a.colind + p);
        const float x = __ldg(a.X + (int64_t)v * F + k);
        if (x > best) { best = x; rec = v; }
      }
      if (slot < 0) {            // the segment is its whole hyperedge
        a.xe[(int64_t)e * F + k] = best * sc;
        a.record[(int64_t)e * F + k] = rec;
      } else {
        atomicMax(a.packed + (int64_t)slot * F + k, pack_max(best, rec));
      }
    }
  }
}

// first segment of every split hyperedge: unpack the winner
__global__ void __launch_bounds__(kThreads) max_finish_kernel(const MArgs a) {
  const int lane = threadIdx.x & 31;
  const int F = a.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarps;
  for (int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); i < a.nheavy_segs; i += nwarps) {
    const int32_t sgm = a.heavy_segs[i];
    const int32_t e = a.seg_edge[sgm];
    if (sgm > 0 && a.seg_edge[sgm - 1] == e) continue;
    const int32_t slot = a.seg_slot[sgm];
    float sc = 1.0f;
    if (a.s1) sc = __ldg(a.s1 + e);
    if (a.s2) sc *= __ldg(a.s2 + e);
    for (int k = lane; k < F; k += 32) {
      float x; int32_t v;
      unpack_max(a.packed[(int64_t)slot * F + k], x, v);
      a.xe[(int64_t)e * F + k] = x * sc;
      a.record[(int64_t)e * F + k] = v;
    }
  }
}

// dX[record[e,k], k] += a_out[record[e,k]] * xe[e,k]   (xe = scale * sum of the members' gradient rows)
__global__ void max_bwd_scatter_kernel(const MArgs a) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= a.nedge * a.F) return;
  const int k = (int)(i % a.F);
  const int32_t v = a.record_in[i];
  const float g = a.xe[i];
  atomicAdd(a.dX + (int64_t)v * a.F + k, a.a_out ? g * __ldg(a.a_out + v) : g);
}

int prologue(const char *what, int64_t num_nodes, int64_t num_edges, const void *indptr,
             const void *indices, const void *in, void *out, int32_t F) {
  HG_REQUIRE(num_nodes >= 0 && num_edges >= 0, "%s: negative size", what);
  HG_REQUIRE(F >= 1, "%s: feature length must be >= 1 (got %d)", what, F);
  HG_REQUIRE(indptr && in && out, "%s: a required pointer is NULL", what);
  HG_REQUIRE(indices || num_edges == 0, "%s: indices is NULL", what);
  return HG_OK;
}

unsigned grid_for(int64_t nedge) {
  int dev = 0, sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = ceil_div<int64_t>(nedge, kWarps), cap = (int64_t)sm * 8;
  return (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

int hg_aggr_mean(int64_t num_nodes, int64_t num_edges, const int32_t *d_t_indptr,
                 const int32_t *d_t_indices, const float *d_X, const float *d_s1, const float *d_s2,
                 const float *d_a_out, float *d_Y, int32_t F, int32_t flags, int device,
                 void *stream) {
  if (int rc = prologue("aggr_mean", num_nodes, num_edges, d_t_indptr, d_t_indices, d_X, d_Y, F)) return rc;
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "aggr_mean: cannot select device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & HG_ACCUMULATE))
    HG_CUDA_TRY(cudaMemsetAsync(d_Y, 0, (size_t)num_nodes * F * sizeof(float), s));
  if (num_edges == 0) return HG_OK;
  VArgs a{};
  a.indptr = d_t_indptr; a.indices = d_t_indices; a.X = d_X; a.s1 = d_s1; a.s2 = d_s2;
  a.a_out = d_a_out; a.Y = d_Y; a.nedge = num_edges; a.F = F;
  edge_kernel<kMean><<<grid_for(num_edges), kThreads, 0, s>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int hg_aggr_max_forward(int64_t num_nodes, int64_t num_edges, const int32_t *d_t_indptr,
                        const int32_t *d_t_indices, const float *d_X, const float *d_s1,
                        const float *d_s2, const float *d_a_out, float *d_Y, int32_t *d_record,
                        int32_t F, int32_t flags, int device, void *stream) {
  if (int rc = prologue("aggr_max_forward", num_nodes, num_edges, d_t_indptr, d_t_indices, d_X, d_Y, F))
    return rc;
  HG_REQUIRE(d_record || num_edges == 0, "aggr_max_forward: record table is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "aggr_max_forward: cannot select device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & HG_ACCUMULATE))
    HG_CUDA_TRY(cudaMemsetAsync(d_Y, 0, (size_t)num_nodes * F * sizeof(float), s));
  if (num_edges == 0) return HG_OK;
  VArgs a{};
  a.indptr = d_t_indptr; a.indices = d_t_indices; a.X = d_X; a.s1 = d_s1; a.s2 = d_s2;
  a.a_out = d_a_out; a.Y = d_Y; a.record_out = d_record; a.nedge = num_edges; a.F = F;
  edge_kernel<kMaxFwd><<<grid_for(num_edges), kThreads, 0, s>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int hg_aggr_max_backward(int64_t num_nodes, int64_t num_edges, const int32_t *d_t_indptr,
                         const int32_t *d_t_indices, const float *d_G, const float *d_s1,
                         const float *d_s2, const float *d_a_out, const int32_t *d_record,
                         float *d_dX, int32_t F, int32_t flags, int device, void *stream) {
  if (int rc = prologue("aggr_max_backward", num_nodes, num_edges, d_t_indptr, d_t_indices, d_G, d_dX, F))
    return rc;
  HG_REQUIRE(d_record || num_edges == 0, "aggr_max_backward: record table is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "aggr_max_backward: cannot select device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & HG_ACCUMULATE))
    HG_CUDA_TRY(cudaMemsetAsync(d_dX, 0, (size_t)num_nodes * F * sizeof(float), s));
  if (num_edges == 0) return HG_OK;
  VArgs a{};
  a.indptr = d_t_indptr; a.indices = d_t_indices; a.G = d_G; a.s1 = d_s1; a.s2 = d_s2;
  a.a_out = d_a_out; a.Y = d_dX; a.record_in = d_record; a.nedge = num_edges; a.F = F;
  edge_kernel<kMaxBwd><<<grid_for(num_edges), kThreads, 0, s>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int hg_plan_max_forward(hgPlan *plan, const int32_t *d_t_indptr, const float *d_X, const float *d_s1,
                        const float *d_s2, const float *d_a_out, float *d_Y, int32_t *d_record, int32_t F,
                        void *stream) {
  HG_REQUIRE(plan != nullptr && d_t_indptr && d_X && d_Y && d_record && F >= 1, "plan_max_forward: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (!plan->canonical || !stream_available(plan, F, true) || F % 4 != 0)   // the reference's own scheme
    return hg_aggr_max_forward(plan->num_nodes, plan->num_edges, d_t_indptr, plan->colind, d_X, d_s1, d_s2, d_a_out,
                               d_Y, d_record, F, 0, plan->device, stream);
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "plan_max_forward: cannot select device %d", plan->device);
  if (int rc = ensure_xe(plan, F, s)) return rc;
  MArgs a{};
  a.key = plan->key; a.colind = plan->colind; a.seg_edge = plan->seg_edge; a.seg_slot = plan->seg_slot;
  a.heavy_segs = plan->heavy_segs; a.X = d_X; a.s1 = d_s1; a.s2 = d_s2; a.xe = plan->xe; a.record = d_record;
  a.nseg = plan->nseg; a.nheavy_segs = plan->nheavy_segs; a.nedge = plan->num_edges; a.F = F;
  if (plan->nheavy_edges > 0) {
    const size_t n = (size_t)plan->nheavy_edges * F;
    if (int rc = plan_grow(plan, &plan->scratch, &plan->scratch_floats, 2 * n, s, "heavy-hyperedge scratch")) return rc;
    a.packed = reinterpret_cast<unsigned long long *>(plan->scratch);
    max_init_kernel<<<(unsigned)ceil_div<int64_t>((int64_t)n, 256), 256, 0, s>>>((int64_t)n, a.packed);
  }
  seg_max_kernel<<<grid_for(plan->nseg), kThreads, 0, s>>>(a);
  if (plan->nheavy_segs > 0) max_finish_kernel<<<grid_for(plan->nheavy_segs), kThreads, 0, s>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  plan->kernels_launched += 1 + (plan->nheavy_segs > 0 ? 2 : 0);
  dev::Args b{};
  b.a_out = d_a_out; b.Y = d_Y; b.F = F; b.X = d_X;
  return launch_stream_stages(plan, b, 2, s);          // Y = a_out . H . Xe, every row written once
}

int hg_plan_max_backward(hgPlan *plan, const int32_t *d_t_indptr, const float *d_G, const float *d_s1,
                         const float *d_s2, const float *d_a_out, const int32_t *d_record, float *d_dX, int32_t F,
                         void *stream) {
  HG_REQUIRE(plan != nullptr && d_t_indptr && d_G && d_dX && d_record && F >= 1, "plan_max_backward: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (!plan->canonical || !stream_available(plan, F, true) || F % 4 != 0)
    return hg_aggr_max_backward(plan->num_nodes, plan->num_edges, d_t_indptr, plan->colind, d_G, d_s1, d_s2, d_a_out,
                                d_record, d_dX, F, 0, plan->device, stream);
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "plan_max_backward: cannot select device %d", plan->device);
  HG_CUDA_TRY(cudaMemsetAsync(d_dX, 0, (size_t)plan->num_nodes * F * sizeof(float), s));
  dev::Args b{};
  b.X = d_G; b.s1 = d_s1; b.s2 = d_s2; b.F = F; b.Y = d_dX;
  if (int rc = launch_stream_stages(plan, b, 1, s)) return rc;     // xe = scale . H^T . G over balancer segments
  MArgs a{};
  a.xe = plan->xe; a.record_in = d_record; a.a_out = d_a_out; a.dX = d_dX; a.nedge = plan->num_edges; a.F = F;
  const int64_t n = plan->num_edges * (int64_t)F;
  max_bwd_scatter_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, s>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  ++plan->kernels_launched;
  return HG_OK;
}

int hg_weight_grad(int64_t num_edges, const int32_t *d_t_indptr, const int32_t *d_t_indices,
                   const float *d_X, const float *d_G, const float *d_s1, const float *d_a_out,
                   const float *d_a_in, float *d_dW, int32_t F, int device, void *stream) {
  if (int rc = prologue("weight_grad", 0, num_edges, d_t_indptr, d_t_indices, d_X, d_dW, F)) return rc;
  HG_REQUIRE(d_G != nullptr, "weight_grad: G is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "weight_grad: cannot select device %d", device);
  if (num_edges == 0) return HG_OK;
  VArgs a{};
  a.indptr = d_t_indptr; a.indices = d_t_indices; a.X = d_X; a.G = d_G; a.s1 = d_s1;
  a.a_out = d_a_out; a.a_in = d_a_in; a.dW = d_dW; a.nedge = num_edges; a.F = F;
  weight_grad_kernel<<<grid_for(num_edges), kThreads, 0, (cudaStream_t)stream>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

}  // extern "C"
