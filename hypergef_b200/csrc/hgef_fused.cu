// hgef_fused.cu -- the single-launch persistent form of the fused aggregation.
//
// Why: on a B200 the two-pass form (cudaMemset Y, then gather/reduce/scatter) moves every Y
// row through DRAM three times -- zero-fill write, read-modify of the vector reductions,
// final write-back -- because Y (hundreds of MB) does not survive in the 126 MB L2 between
// the memset and the kernel.  ncu on the two-pass kernel: 2.08 GB of DRAM traffic + 0.65 GB of
// memset for 1.31 GB of algorithmic bytes (profiles/).  Here the zero-fill happens INSIDE the
// kernel, a few microseconds before the first reduction reaches the row, so the zeros are
// still dirty lines in L2 when `red` hits them and a Y row goes to DRAM exactly once.
//
// Schedule: CTAs claim TILES of consecutive segments from a global counter (in order).
//   phase 0  the CTA stages the tile's slice of the flagged column indices, segment bounds,
//            slots and scales in shared memory (coalesced), and zero-fills the Y rows whose
//            FIRST occurrence (in H_T_colind order) lies in the tile -- plus its share of
//            the vertices that are in no hyperedge; then publishes done[tile] and bumps the
//            finished-tile count of its block of 32 tiles (release).
//   phase 1  warps take segments of the tile from a shared cursor; gather + reduce in
//            registers (indices come from shared memory, so the X loads issue immediately).
//   phase 2  before its first scatter of the tile a warp waits until every tile <= its own has
//            finished phase 0 (every row it can touch was first-touched by one of those); by
//            then that is almost always already true, the wait hides behind the gather.  Rows with a single
//            occurrence in the whole graph are written with plain 128-bit stores (no
//            zero-fill, no reduction); the rest use red.global.add.v4.f32.
// Deadlock freedom: a tile id only exists once a RUNNING CTA has claimed it, claims are in
// order, and phase 0 never waits -- so every tile a waiter depends on completes.
// Heavy hyperedges (w > 1 segments) publish partial sums to scratch here; the second,
// small launch (seg_pass2 with the same flags) scatters them.
#include <cstdlib>

#include "hgef_aggr.cuh"

namespace hg {
namespace {
using namespace dev;

constexpr uint32_t kFirst = 0x80000000u, kExcl = 0x40000000u, kIdMask = 0x3fffffffu;
constexpr int kIdxCap = 3072;  // staged column indices per tile (ints); beyond that: global reads

struct FusedArgs {
  Args a;
  const int32_t *cflag, *iso;
  int32_t *ctrl;        // [0] tile counter, [2] give-up flag, [8 + b] #finished tiles of block b
                        // (32 tiles per block), [8 + nblk + t] done flag of tile t
  int64_t niso;
  int32_t ntiles, tile_segs, nblk;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_inc(int *p) {
  asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}

// Wait (whole warp) until the zero-fill of every tile <= t is complete.  Progress is tracked in
// blocks of 32 tiles: blk_cnt[b] == 32 means block b is done, and one warp-wide load checks 32
// blocks (1024 tiles), so the per-warp watermark `blk_wm` catches up in a few L2 round trips over
// the whole kernel; the tiles of t's own block are checked flag by flag.  Normally everything is
// already complete (tiles are claimed in order and zero-fill first); the spin is bounded so that a
// protocol bug cannot hang the GPU -- ctrl[2] records the give-up and hg_plan_check reports it.
__device__ __forceinline__ void wait_zero_fill(int32_t *ctrl, int nblk, int t, int lane, int &blk_wm) {
  const int *blk_cnt = ctrl + 8, *done = ctrl + 8 + nblk;
  const int need = t >> 5;
  unsigned spins = 0;
  // polls are relaxed (an acquire load costs an L1 invalidate each time); one acquire fence at the end
  while (blk_wm < need) {
    const int b = blk_wm + lane;
    const int c = b < need ? ld_relaxed(blk_cnt + b) : 32;
    const unsigned full = __ballot_sync(kFull, c == 32);
    const int lead = full == kFull ? 32 : __ffs(~full) - 1;
    blk_wm = min(blk_wm + lead, need);
    if (lead < 32 && blk_wm < need) {
      __nanosleep(32);
      if (++spins > (1u << 20)) { if (lane == 0) atomicExch(ctrl + 2, 1); break; }
    }
  }
  const int base = need << 5;
  for (;;) {
    const int d = base + lane <= t ? ld_relaxed(done + base + lane) : 1;
    if (__all_sync(kFull, d == 1)) break;
    __nanosleep(32);
    if (++spins > (1u << 20)) { if (lane == 0) atomicExch(ctrl + 2, 1); break; }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

__device__ __forceinline__ void st_zero_v4(float *p) {
  asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(0.0f) : "memory");
}
__device__ __forceinline__ void st_v4(float *p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

template <int VPL>
__global__ void __launch_bounds__(kThreads) fused_kernel(const FusedArgs fa) {
  extern __shared__ int32_t smem[];
  const Args &a = fa.a;
  const int T = fa.tile_segs;
  const int buf_ints = 3 * T + 1 + kIdxCap;  // per tile buffer: key[T+1] slot[T] scale[T] idx[kIdxCap]
  __shared__ int s_tile[2], s_cursor;

  const int tid = threadIdx.x, lane = tid & 31;
  const int lpr = a.lpr, groups = 32 / lpr, grp = lane / lpr;
  const int col = (lane & (lpr - 1)) * 4;
  const int F = a.F;
  const int64_t S = a.nwork;
  int *counter = fa.ctrl, *blk_cnt = fa.ctrl + 8, *done = fa.ctrl + 8 + fa.nblk;
  int blk_wm = 0;  // per warp: blocks [0, blk_wm) are known to be completely zero-filled

  // phase 0 of tile `t` into buffer `b`: stage, zero-fill first-touched rows, publish.
  auto phase0 = [&](int t, int b) {
    int32_t *s_key = smem + b * buf_ints, *s_slot = s_key + T + 1;
    float *s_scale = reinterpret_cast<float *>(s_slot + T);
    int32_t *s_idx = reinterpret_cast<int32_t *>(s_scale + T);
    const int64_t s0 = (int64_t)t * T;
    const int nseg = (int)min((int64_t)T, S - s0);
    const int32_t p0 = __ldg(a.key + s0), p1 = __ldg(a.key + s0 + nseg);
    for (int i = tid; i <= nseg; i += kThreads) s_key[i] = __ldg(a.key + s0 + i);
    for (int i = tid; i < nseg; i += kThreads) {
      s_slot[i] = __ldg(a.seg_slot + s0 + i);
      s_scale[i] = edge_scale(a, __ldg(a.seg_edge + s0 + i));
    }
    const int nidx = min(p1 - p0, kIdxCap);
    for (int i = tid; i < nidx; i += kThreads) s_idx[i] = __ldg(fa.cflag + p0 + i);
    __syncthreads();
    const int rows_per_pass = kThreads / lpr;
    const int rgrp = tid / lpr;
    for (int32_t pb = p0; pb < p1; pb += rows_per_pass) {
      const int32_t p = pb + rgrp;
      if (p < p1) {
        const uint32_t c = (uint32_t)(p - p0 < kIdxCap ? s_idx[p - p0] : __ldg(fa.cflag + p));
        if ((c & kFirst) && !(c & kExcl)) {
          float *yp = a.Y + (int64_t)(c & kIdMask) * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (col + j * 128 < F) st_zero_v4(yp + j * 128);
        }
      }
    }
    // this tile's share of the vertices that no hyperedge touches
    const int64_t i0 = fa.niso * t / fa.ntiles, i1 = fa.niso * (t + 1) / fa.ntiles;
    for (int64_t i = i0 + rgrp; i < i1; i += rows_per_pass) {
      float *yp = a.Y + (int64_t)__ldg(fa.iso + i) * F + col;
#pragma unroll
      for (int j = 0; j < VPL; ++j)
        if (col + j * 128 < F) st_zero_v4(yp + j * 128);
    }
    __syncthreads();  // all zero stores of the CTA are ordered before the release below
    if (tid == 0) {
      st_release(done + t, 1);
      red_release_inc(blk_cnt + (t >> 5));
    }
  };

  // A claimed tile starts its zero-fill at once (every later tile's scatter waits for it), but
  // its segments are processed one tile later: by then the zero-fills of all earlier tiles,
  // which were claimed before it, have long been published and the wait below is free.
  if (tid == 0) s_tile[0] = atomicAdd(counter, 1);
  __syncthreads();
  int t = s_tile[0], cb = 0;
  if (t >= fa.ntiles) return;
  phase0(t, 0);
  for (;;) {
    __syncthreads();  // everyone is done with the tile that lived in buffer cb ^ 1
    if (tid == 0) {
      s_tile[cb ^ 1] = atomicAdd(counter, 1);
      s_cursor = 0;
    }
    __syncthreads();
    const int t_next = s_tile[cb ^ 1];
    if (t_next < fa.ntiles) phase0(t_next, cb ^ 1);

    const int32_t *s_key = smem + cb * buf_ints, *s_slot = s_key + T + 1;
    const float *s_scale = reinterpret_cast<const float *>(s_slot + T);
    const int32_t *s_idx = reinterpret_cast<const int32_t *>(s_scale + T);
    const int nseg = (int)min((int64_t)T, S - (int64_t)t * T);
    const int32_t p0 = s_key[0];

    // ---------------- phases 1 + 2: segments of the tile, one warp each ----------------
    bool may_scatter = false;
    for (;;) {
      int i = 0;
      if (lane == 0) i = atomicAdd(&s_cursor, 1);
      i = __shfl_sync(kFull, i, 0);
      if (i >= nseg) break;
      const int32_t lo = s_key[i], hi = s_key[i + 1];
      const int32_t slot = s_slot[i];
      Acc<VPL> acc;
      acc.zero();
      for (int32_t base = lo; base < hi; base += 32) {
        const int n = min(32, hi - base);
        uint32_t my_c = 0;
        float my_a = 1.0f;
        if (lane < n) {
          const int32_t p = base + lane;
          my_c = (uint32_t)(p - p0 < kIdxCap ? s_idx[p - p0] : __ldg(fa.cflag + p));
          if (a.a_in) my_a = __ldg(a.a_in + (my_c & kIdMask));
        }
#pragma unroll 4
        for (int r0 = 0; r0 < n; r0 += groups) {
          const int r = r0 + grp;
          const uint32_t c = __shfl_sync(kFull, my_c, r & 31);
          const float w = __shfl_sync(kFull, my_a, r & 31);
          if (r < n) {
            const float *xp = a.X + (int64_t)(c & kIdMask) * F + col;
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
      if (slot >= 0) {  // heavy hyperedge: publish the partial sum, pass 2 scatters
        if (lane < lpr) {
          float *sp = a.scratch + (int64_t)slot * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (col + j * 128 < F) red_add_v4(sp + j * 128, acc.v[j]);
        }
        continue;
      }
      scale_acc<VPL>(acc, s_scale[i]);
      if (!may_scatter) {  // every row this tile can touch was first-touched by a tile <= t
        wait_zero_fill(fa.ctrl, fa.nblk, t, lane, blk_wm);
        __syncwarp();
        may_scatter = true;
      }
      for (int32_t base = lo; base < hi; base += 32) {
        const int n = min(32, hi - base);
        uint32_t my_c = 0;
        float my_o = 1.0f;
        if (lane < n) {
          const int32_t p = base + lane;
          my_c = (uint32_t)(p - p0 < kIdxCap ? s_idx[p - p0] : __ldg(fa.cflag + p));
          if (a.a_out) my_o = __ldg(a.a_out + (my_c & kIdMask));
        }
#pragma unroll 4
        for (int r0 = 0; r0 < n; r0 += groups) {
          const int r = r0 + grp;
          const uint32_t c = __shfl_sync(kFull, my_c, r & 31);
          const float o = __shfl_sync(kFull, my_o, r & 31);
          if (r < n) {
            float *yp = a.Y + (int64_t)(c & kIdMask) * F + col;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
              if (col + j * 128 < F) {
                const float4 v = make_float4(acc.v[j].x * o, acc.v[j].y * o, acc.v[j].z * o, acc.v[j].w * o);
                if (c & kExcl) st_v4(yp + j * 128, v);
                else red_add_v4(yp + j * 128, v);
              }
            }
          }
        }
      }
    }
    if (t_next >= fa.ntiles) break;
    t = t_next;
    cb ^= 1;
  }
}

// Pass 2 for heavy hyperedges with the single-writer flags (rows that were never zero-filled
// must be stored, not reduced).
template <int VPL>
__global__ void __launch_bounds__(kThreads) fused_pass2_kernel(const FusedArgs fa) {
  const Args &a = fa.a;
  const int lane = threadIdx.x & 31;
  const int lpr = a.lpr, groups = 32 / lpr, grp = lane / lpr;
  const int col = (lane & (lpr - 1)) * 4;
  const int F = a.F;
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); i < a.nwork; i += nwarps) {
    const int32_t s = __ldg(a.seg_list + i);
    const int32_t lo = __ldg(a.key + s), hi = __ldg(a.key + s + 1);
    const float *sp = a.scratch + (int64_t)__ldg(a.seg_slot + s) * F + col;
    const float sc = edge_scale(a, __ldg(a.seg_edge + s));
    Acc<VPL> acc;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      acc.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + j * 128 < F) acc.v[j] = *reinterpret_cast<const float4 *>(sp + j * 128);
    }
    scale_acc<VPL>(acc, sc);
    for (int32_t base = lo; base < hi; base += 32) {
      const int n = min(32, hi - base);
      uint32_t my_c = 0;
      float my_o = 1.0f;
      if (lane < n) {
        my_c = (uint32_t)__ldg(fa.cflag + base + lane);
        if (a.a_out) my_o = __ldg(a.a_out + (my_c & kIdMask));
      }
#pragma unroll 4
      for (int r0 = 0; r0 < n; r0 += groups) {
        const int r = r0 + grp;
        const uint32_t c = __shfl_sync(kFull, my_c, r & 31);
        const float o = __shfl_sync(kFull, my_o, r & 31);
        if (r < n) {
          float *yp = a.Y + (int64_t)(c & kIdMask) * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j) {
            if (col + j * 128 < F) {
              const float4 v = make_float4(acc.v[j].x * o, acc.v[j].y * o, acc.v[j].z * o, acc.v[j].w * o);
              if (c & kExcl) st_v4(yp + j * 128, v);
              else red_add_v4(yp + j * 128, v);
            }
          }
        }
      }
    }
  }
}

template <int VPL>
int launch(hgPlan *plan, FusedArgs &fa, cudaStream_t s) {
  const int F = fa.a.F;
  // tile size: the tiles in flight (2 per resident CTA) plus the rows still waiting for their
  // last reduction must stay L2-resident, so tiles are small: ~64 KB of gathered + scattered rows
  const double avg_len = (double)plan->nnz / (double)plan->nseg;
  static const double tile_kb = getenv("HGEF_TILE_KB") ? atof(getenv("HGEF_TILE_KB")) : 64.0;
  int T = (int)(tile_kb * 1024.0 / (avg_len * 4.0 * F * 2.0));
  T = T < 8 ? 8 : (T > 512 ? 512 : T);
  T = (T + 7) & ~7;
  fa.tile_segs = T;
  fa.ntiles = (int32_t)ceil_div<int64_t>(plan->nseg, T);
  const size_t smem = (size_t)2 * (3 * T + 1 + kIdxCap) * sizeof(int32_t);
  static bool attr_set[8] = {};
  if (!attr_set[VPL]) {
    HG_CUDA_TRY(cudaFuncSetAttribute(fused_kernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_set[VPL] = true;
  }
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_kernel<VPL>, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)plan->sm_count * per_sm;
  if (grid > fa.ntiles) grid = fa.ntiles;
  fa.nblk = (fa.ntiles + 31) / 32;
  HG_CUDA_TRY(cudaMemsetAsync(plan->ctrl, 0, (size_t)(fa.ntiles + fa.nblk + 8) * sizeof(int32_t), s));
  fa.ctrl = plan->ctrl;
  fa.a.nwork = plan->nseg;
  fused_kernel<VPL><<<(unsigned)grid, kThreads, smem, s>>>(fa);
  HG_CUDA_TRY(cudaGetLastError());
  if (plan->nheavy_segs > 0) {
    fa.a.nwork = plan->nheavy_segs;
    fa.a.seg_list = plan->heavy_segs;
    int64_t g2 = ceil_div<int64_t>(plan->nheavy_segs, kWarpsPerBlock);
    if (g2 > (int64_t)plan->sm_count * 8) g2 = (int64_t)plan->sm_count * 8;
    fused_pass2_kernel<VPL><<<(unsigned)g2, kThreads, 0, s>>>(fa);
    HG_CUDA_TRY(cudaGetLastError());
  }
  return HG_OK;
}

}  // namespace

int fused_check(hgPlan *plan, cudaStream_t s) {
  if (!plan->ctrl) return HG_OK;
  int32_t stalled = 0;
  HG_CUDA_TRY(cudaMemcpyAsync(&stalled, plan->ctrl + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  if (stalled)
    return set_error(HG_ECUDA, "fused aggregation: a tile gave up waiting for the zero-fill watermark; "
                               "the last result is invalid");
  return HG_OK;
}

bool fused_available(const hgPlan *plan) { return plan->cflag != nullptr && plan->ctrl != nullptr; }

// Y is NOT zero-filled by the caller; scratch (if any) is.
int launch_fused(hgPlan *plan, const dev::Args &base, cudaStream_t s) {
  FusedArgs fa{};
  fa.a = base;
  fa.cflag = plan->cflag;
  fa.iso = plan->iso_list;
  fa.niso = plan->niso;
  if (base.F <= 128) return launch<1>(plan, fa, s);
  if (base.F <= 256) return launch<2>(plan, fa, s);
  return launch<4>(plan, fa, s);
}

}  // namespace hg
