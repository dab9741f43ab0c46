// hgef_aggr.cu -- the fused two-stage aggregation for sm_100a.
//
//     Y = diag(a_out) . H . diag(s1*s2) . H^T . diag(a_in) . X
//
// One warp owns one balancer SEGMENT at a time (a slice of <= ngs members of one hyperedge):
//   stage 1  gather the member rows of X with 128-bit read-only loads -- a feature row is
//            contiguous, F/4 lanes cover it and 32/(F/4) rows are in flight per warp step --
//            and reduce them in registers, finishing with a shuffle butterfly across the
//            row groups: a warp-level segmented reduction, no atomics;
//   scale    by s1[e]*s2[e] (degE, W);
//   stage 2  scatter acc * a_out[v] to the same member rows of Y with 128-bit vector
//            reductions (red.global.add.v4.f32, one L2 transaction per 16 B instead of
//            the reference's four scalar atomics).
// The hyperedge feature lives in registers only.  A hyperedge that the balancer cut into
// w > 1 segments ("heavy") reduces its w partial sums into one L2-resident scratch row
// and a second, small launch scatters the completed feature over the same segments, so
// every member row is still gathered exactly once (the reference gathers it w times).
//
// The literal group schedule of the reference (hgnnaggr_cuda.cu:14-47) is kept as
// hg_aggr_groups: same device code, one warp per balancer group.
#include "hgef_stream.cuh"

namespace hg {
#ifdef HGEF_LAB
bool fused_available(const hgPlan *plan, int F, bool force);
int launch_fused(hgPlan *plan, const dev::Args &base, cudaStream_t s);
int fused_check(hgPlan *plan, cudaStream_t s);
bool pull_available(const hgPlan *plan, int F, bool force);
int launch_pull_any(hgPlan *plan, const dev::Args &base, cudaStream_t s);
#endif
bool stream_available(const hgPlan *plan, int F, bool force);
int launch_stream(hgPlan *plan, const dev::Args &base, cudaStream_t s);
int launch_stream_padded(hgPlan *plan, const dev::Args &base, cudaStream_t s);
namespace {
using namespace dev;

// Segment schedule, pass 1: all segments.  Light hyperedges finish here.
template <int VPL>
__global__ void __launch_bounds__(kThreads) seg_pass1_kernel(const Args a) {
  const int lane = threadIdx.x & 31;
  const int col = (lane & (a.lpr - 1)) * 4;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t s = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); s < a.nwork; s += nwarps) {
    const int32_t lo = __ldg(a.key + s), hi = __ldg(a.key + s + 1);
    const int32_t slot = __ldg(a.seg_slot + s);
    Acc<VPL> acc;
    acc.zero();
    gather_rows<VPL>(a, lo, hi, lane, col, acc);
    if (slot < 0) {
      scale_acc<VPL>(acc, edge_scale(a, __ldg(a.seg_edge + s)));
      scatter_rows<VPL>(a, lo, hi, lane, col, acc);
    } else if (lane < a.lpr) {  // one row group publishes the partial sum
      float *sp = a.scratch + (int64_t)slot * a.F + col;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (col + j * 128 < a.F) red_add_v4(sp + j * 128, acc.v[j]);
    }
  }
}

// Segment schedule, pass 2: the segments of heavy hyperedges scatter the completed feature.
template <int VPL>
__global__ void __launch_bounds__(kThreads) seg_pass2_kernel(const Args a) {
  const int lane = threadIdx.x & 31;
  const int col = (lane & (a.lpr - 1)) * 4;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < a.nwork; i += nwarps) {
    const int32_t s = __ldg(a.seg_list + i);
    const int32_t lo = __ldg(a.key + s), hi = __ldg(a.key + s + 1);
    const float *sp = a.scratch + (int64_t)__ldg(a.seg_slot + s) * a.F + col;
    const float sc = edge_scale(a, __ldg(a.seg_edge + s));
    Acc<VPL> acc;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      acc.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      // plain (coherent) load: the scratch row was written by pass 1 of this call
      if (col + j * 128 < a.F) acc.v[j] = *reinterpret_cast<const float4 *>(sp + j * 128);
    }
    scale_acc<VPL>(acc, sc);
    scatter_rows<VPL>(a, lo, hi, lane, col, acc);
  }
}

// Reference schedule: one warp per balancer group (read seg st[g], write seg ed[g]).
template <int VPL>
__global__ void __launch_bounds__(kThreads) group_kernel(const Args a) {
  const int lane = threadIdx.x & 31;
  const int col = (lane & (a.lpr - 1)) * 4;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t g = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); g < a.nwork; g += nwarps) {
    const int32_t rs = __ldg(a.st + g), ws = __ldg(a.ed + g);
    Acc<VPL> acc;
    acc.zero();
    gather_rows<VPL>(a, __ldg(a.key + rs), __ldg(a.key + rs + 1), lane, col, acc);
    scale_acc<VPL>(acc, edge_scale(a, __ldg(a.row + g)));
    scatter_rows<VPL>(a, __ldg(a.key + ws), __ldg(a.key + ws + 1), lane, col, acc);
  }
}

// ---- scalar path: any F (lane l holds columns l, l+32, ...) -- scalar atomics like the
// reference; used when F % 4 != 0 (e.g. the 7-class output layer) or pointers are unaligned.
enum { kModeSeg1 = 0, kModeSeg2 = 1, kModeGroup = 2 };

template <int MODE>
__global__ void __launch_bounds__(kThreads) scalar_kernel(const Args a) {
  const int lane = threadIdx.x & 31;
  const int F = a.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < a.nwork; i += nwarps) {
    int32_t rlo, rhi, wlo, whi, e, slot = -1;
    if (MODE == kModeGroup) {
      const int32_t rs = a.st[i], ws = a.ed[i];
      rlo = a.key[rs]; rhi = a.key[rs + 1]; wlo = a.key[ws]; whi = a.key[ws + 1]; e = a.row[i];
    } else {
      const int32_t s = MODE == kModeSeg2 ? a.seg_list[i] : (int32_t)i;
      rlo = wlo = a.key[s]; rhi = whi = a.key[s + 1]; e = a.seg_edge[s]; slot = a.seg_slot[s];
    }
    const float sc = edge_scale(a, e);
    for (int k = lane; k < F; k += 32) {
      float acc = 0.f;
      if (MODE == kModeSeg2) {
        acc = a.scratch[(int64_t)slot * F + k];
      } else {
        for (int32_t p = rlo; p < rhi; ++p) {
          const int32_t v = __ldg(a.colind + p);
          const float x = __ldg(a.X + (int64_t)v * F + k);
          acc = a.a_in ? fmaf(__ldg(a.a_in + v), x, acc) : acc + x;
        }
      }
      if (MODE == kModeSeg1 && slot >= 0) {
        atomicAdd(a.scratch + (int64_t)slot * F + k, acc);
        continue;
      }
      acc *= sc;
      for (int32_t p = wlo; p < whi; ++p) {
        const int32_t v = __ldg(a.colind + p);
        atomicAdd(a.Y + (int64_t)v * F + k, a.a_out ? acc * __ldg(a.a_out + v) : acc);
      }
    }
  }
}

// ---------------------------------------------------------------- host side
inline int lanes_per_row(int F) {  // power of two >= min(F,128)/4
  int need = (F < 128 ? F : 128) / 4, l = 1;
  while (l < need) l <<= 1;
  return l;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline unsigned grid_for(int64_t nwork, int sm_count, int blocks_per_sm) {
  int64_t need = ceil_div<int64_t>(nwork, kWarpsPerBlock);
  int64_t cap = (int64_t)sm_count * blocks_per_sm;
  return (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
}

#define HG_DISPATCH_VPL(F, KERNEL, grid, stream, args)                        \
  do {                                                                        \
    if ((F) <= 128) KERNEL<1><<<grid, kThreads, 0, stream>>>(args);           \
    else if ((F) <= 256) KERNEL<2><<<grid, kThreads, 0, stream>>>(args);      \
    else KERNEL<4><<<grid, kThreads, 0, stream>>>(args);                      \
  } while (0)

int check_common(const float *X, float *Y, int32_t F) {
  HG_REQUIRE(X != nullptr && Y != nullptr, "aggr: X or Y is NULL");
  HG_REQUIRE(F >= 1, "aggr: feature length must be >= 1 (got %d)", F);
  return HG_OK;
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

int hg_aggr_groups(int64_t num_nodes, int64_t ngroup, const int32_t *d_key, const int32_t *d_row,
                   const int32_t *d_st, const int32_t *d_ed, const int32_t *d_t_indices,
                   const float *d_X, const float *d_s1, const float *d_s2, const float *d_a_out,
                   const float *d_a_in, float *d_Y, int32_t F, int32_t flags, int device,
                   void *stream) {
  if (int rc = check_common(d_X, d_Y, F)) return rc;
  HG_REQUIRE(num_nodes >= 0 && ngroup >= 0, "aggr_groups: negative size");
  HG_REQUIRE(ngroup == 0 || (d_key && d_row && d_st && d_ed && d_t_indices),
             "aggr_groups: a schedule array is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "aggr_groups: cannot select device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  if (!(flags & HG_ACCUMULATE))
    HG_CUDA_TRY(cudaMemsetAsync(d_Y, 0, (size_t)num_nodes * F * sizeof(float), s));
  if (ngroup == 0) return HG_OK;
  int sm = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  Args a{};
  a.key = d_key; a.colind = d_t_indices; a.row = d_row; a.st = d_st; a.ed = d_ed;
  a.X = d_X; a.s1 = d_s1; a.s2 = d_s2; a.a_out = d_a_out; a.a_in = d_a_in; a.Y = d_Y;
  a.nwork = ngroup; a.F = F; a.lpr = lanes_per_row(F);
  const bool vec = F % 4 == 0 && F <= 512 && !(flags & HG_FORCE_SCALAR) && aligned16(d_X) && aligned16(d_Y);
  const unsigned grid = grid_for(ngroup, sm, 8);
  if (vec) HG_DISPATCH_VPL(F, group_kernel, grid, s, a);
  else scalar_kernel<kModeGroup><<<grid, kThreads, 0, s>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int hg_plan_launches(const hgPlan *plan, int64_t *kernels) {
  HG_REQUIRE(plan != nullptr && kernels != nullptr, "plan_launches: NULL argument");
  *kernels = plan->kernels_launched;
  return HG_OK;
}

int hg_plan_debug(hgPlan *plan, int32_t *h_out8, void *stream) {
  HG_REQUIRE(plan != nullptr && h_out8 != nullptr, "plan_debug: NULL argument");
  for (int i = 0; i < 8; ++i) h_out8[i] = 0;
#ifdef HGEF_LAB
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "plan_debug: cannot select device %d", plan->device);
  return ring_debug(plan, h_out8, (cudaStream_t)stream);
#else
  (void)stream;
  return HG_OK;
#endif
}

int hg_plan_check(hgPlan *plan, void *stream) {
  HG_REQUIRE(plan != nullptr, "plan_check: plan is NULL");
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "plan_check: cannot select device %d", plan->device);
  cudaStream_t s = (cudaStream_t)stream;
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  HG_CUDA_TRY(cudaGetLastError());
#ifdef HGEF_LAB
  if (int rc = ring_check(plan, s)) return rc;
  if (int rc = fused_check(plan, s)) return rc;
#endif
  return HG_OK;
}

int hg_plan_reserve(hgPlan *plan, int32_t F_max, void *stream) {
  HG_REQUIRE(plan != nullptr && F_max >= 1, "plan_reserve: bad argument");
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "plan_reserve: cannot select device %d", plan->device);
  cudaStream_t s = (cudaStream_t)stream;
  const int Fp = (F_max + 3) / 4 * 4;
  if (plan->canonical && plan->st_ready)
    if (int rc = plan_grow(plan, &plan->xe, &plan->xe_floats, (size_t)plan->num_edges * Fp, s, "hyperedge features")) return rc;
  if (plan->nheavy_edges > 0)
    if (int rc = plan_grow(plan, &plan->scratch, &plan->scratch_floats, (size_t)plan->nheavy_edges * Fp, s, "heavy-hyperedge scratch")) return rc;
  return HG_OK;
}

// The two stages on their own (include/hgef_b200.h): the balanced stream kernels with the hyperedge features in a
// caller-owned buffer, so that work can be put between them (a projection of the E hyperedge rows instead of the N
// vertex rows -- SURVEY 8(f) N1 -- or an exchange).
static int stage_common(hgPlan *plan, const void *in, void *out, int32_t F, const char *what) {
  HG_REQUIRE(plan != nullptr, "%s: plan is NULL", what);
  HG_REQUIRE(in != nullptr && out != nullptr, "%s: a feature pointer is NULL", what);
  HG_REQUIRE(F >= 4 && F % 4 == 0, "%s: the feature length must be a multiple of 4 (got %d); pad the rows", what, F);
  HG_REQUIRE(aligned16(in) && aligned16(out), "%s: feature pointers must be 16-byte aligned", what);
  HG_REQUIRE(plan->canonical && plan->st_ready, "%s: the plan has no row programs (the schedule is not the balancer's "
             "canonical cross product); use hg_edge_reduce / hg_edge_scatter on the CSR instead", what);
  return HG_OK;
}

int hg_plan_edge_reduce(hgPlan *plan, const float *d_X, const float *d_s1, const float *d_s2, const float *d_a_in,
                        float *d_Xe, int32_t F, void *stream) {
  if (int rc = stage_common(plan, d_X, d_Xe, F, "plan_edge_reduce")) return rc;
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "plan_edge_reduce: cannot select device %d", plan->device);
  Args a{};
  a.X = d_X; a.s1 = d_s1; a.s2 = d_s2; a.a_in = d_a_in; a.F = F;
  return launch_stream_stages(plan, a, 1, (cudaStream_t)stream, d_Xe);
}

int hg_plan_edge_scatter(hgPlan *plan, const float *d_Xe, const float *d_a_out, float *d_Y, int32_t F, void *stream) {
  if (int rc = stage_common(plan, d_Xe, d_Y, F, "plan_edge_scatter")) return rc;
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "plan_edge_scatter: cannot select device %d", plan->device);
  Args a{};
  a.a_out = d_a_out; a.Y = d_Y; a.F = F;
  return launch_stream_stages(plan, a, 2, (cudaStream_t)stream, const_cast<float *>(d_Xe));
}

int hg_aggr_forward(hgPlan *plan, const float *d_X, const float *d_s1, const float *d_s2,
                    const float *d_a_out, const float *d_a_in, float *d_Y, int32_t F, int32_t flags,
                    void *stream) {
  HG_REQUIRE(plan != nullptr, "aggr_forward: plan is NULL");
  if (int rc = check_common(d_X, d_Y, F)) return rc;
  constexpr int kLabForms = HG_FORCE_FUSED | HG_FORCE_PULL | HG_FORCE_RING | HG_FORCE_FSTREAM;
#ifndef HGEF_LAB
  HG_REQUIRE(!(flags & kLabForms), "aggr_forward: flags 0x%x select an experimental kernel form; those are only built "
             "into libhgef_b200_lab.so (make -C hypergef_b200/csrc lab)", flags);
#endif
  if (!plan->canonical)
    return hg_aggr_groups(plan->num_nodes, plan->ngroup, plan->key, plan->row, plan->st, plan->ed,
                          plan->colind, d_X, d_s1, d_s2, d_a_out, d_a_in, d_Y, F, flags, plan->device,
                          stream);
  DeviceGuard guard(plan->device);
  HG_REQUIRE(guard.ok(), "aggr_forward: cannot select device %d", plan->device);
  cudaStream_t s = (cudaStream_t)stream;
  const bool vec4 = F % 4 == 0 && !(flags & HG_FORCE_SCALAR) && aligned16(d_X) && aligned16(d_Y);
  const bool vec = vec4 && F <= 512;
  const bool plain = !(flags & (HG_ACCUMULATE | HG_TWO_PASS | kLabForms));
  Args pa{};
  pa.X = d_X; pa.s1 = d_s1; pa.s2 = d_s2; pa.a_out = d_a_out; pa.a_in = d_a_in; pa.Y = d_Y; pa.F = F;
#ifdef HGEF_LAB
  // experimental forms (hgef_fstream.cu, hgef_ring.cu, hgef_fused.cu): only when asked for
  if (vec4 && (flags & HG_FORCE_FSTREAM) && fstream_available(plan, F, true)) return launch_fstream(plan, pa, s);
  if (vec4 && (flags & HG_FORCE_RING) && ring_available(plan, F, true)) return launch_ring(plan, pa, s);
  if (vec && (flags & HG_FORCE_PULL) && pull_available(plan, F, true)) {
    plan->kernels_launched += 2 + (plan->nheavy_segs > 0 ? 1 : 0);
    return launch_pull_any(plan, pa, s);
  }
#endif
  // stream form: both stages as lean register-only row streams, two launches, the hyperedge features make one
  // round trip through the L2 / HBM (hgef_stream.cu): the form chosen when Y exceeds the L2
  if (vec4 && (plain || (flags & HG_FORCE_STREAM)) && !(flags & (HG_ACCUMULATE | HG_TWO_PASS)) &&
      stream_available(plan, F, (flags & HG_FORCE_STREAM) != 0))
    return launch_stream(plan, pa, s);
  // a feature length that is not a multiple of 4 on a graph that large: the same kernels on rows padded to the
  // next multiple of 4 (one extra pass over X and Y) instead of scalar atomics over a zero-filled Y
  if (F % 4 != 0 && plain && !(flags & HG_FORCE_SCALAR) && stream_available(plan, (F + 3) / 4 * 4, false))
    return launch_stream_padded(plan, pa, s);
#ifdef HGEF_LAB
  const bool fused = vec && (flags & HG_FORCE_FUSED) && !(flags & HG_ACCUMULATE) && fused_available(plan, F, true);
#else
  const bool fused = false;
#endif
  // two-pass form: memset + one warp per balancer segment with vector reductions (small graphs: Y is L2-resident)
  if (!fused && !(flags & HG_ACCUMULATE))
    HG_CUDA_TRY(cudaMemsetAsync(d_Y, 0, (size_t)plan->num_nodes * F * sizeof(float), s));
  const bool heavy = plan->nheavy_segs > 0;
  if (heavy) {
    const size_t need = (size_t)plan->nheavy_edges * F;
    if (int rc = plan_grow(plan, &plan->scratch, &plan->scratch_floats, need, s, "heavy-hyperedge scratch")) return rc;
    HG_CUDA_TRY(cudaMemsetAsync(plan->scratch, 0, need * sizeof(float), s));
  }
  Args a{};
  a.key = plan->key; a.colind = plan->colind; a.seg_edge = plan->seg_edge; a.seg_slot = plan->seg_slot;
  a.X = d_X; a.s1 = d_s1; a.s2 = d_s2; a.a_out = d_a_out; a.a_in = d_a_in; a.Y = d_Y;
  a.scratch = plan->scratch; a.F = F; a.lpr = lanes_per_row(F);
  plan->kernels_launched += 1 + (plan->nheavy_segs > 0 ? 1 : 0);
#ifdef HGEF_LAB
  if (fused) return launch_fused(plan, a, s);
#endif
  a.nwork = plan->nseg;
  unsigned grid = grid_for(plan->nseg, plan->sm_count, 8);
  if (vec) HG_DISPATCH_VPL(F, seg_pass1_kernel, grid, s, a);
  else scalar_kernel<kModeSeg1><<<grid, kThreads, 0, s>>>(a);
  HG_CUDA_TRY(cudaGetLastError());
  if (heavy) {
    a.nwork = plan->nheavy_segs;
    a.seg_list = plan->heavy_segs;
    grid = grid_for(plan->nheavy_segs, plan->sm_count, 8);
    if (vec) HG_DISPATCH_VPL(F, seg_pass2_kernel, grid, s, a);
    else scalar_kernel<kModeSeg2><<<grid, kThreads, 0, s>>>(a);
    HG_CUDA_TRY(cudaGetLastError());
  }
  return HG_OK;
}

}  // extern "C"
