// hgef_fused.cu -- the large-graph forms of the fused aggregation (Y does not fit the L2).
//
// On a B200 the two-pass form of hgef_aggr.cu (cudaMemset Y, then gather / reduce / scatter) moves
// every Y row through DRAM three times -- zero-fill write, read-modify of the vector reductions,
// final write-back -- because Y (hundreds of MB) does not survive in the 126 MB L2 between the
// memset and the kernel: ncu measured 2.08 GB + 0.65 GB (memset) of DRAM traffic for 1.31 GB of
// algorithmic bytes.  This file holds the two forms that replace it, plus one kept for A/B:
//
//   fused_kernel   persistent SCATTER form (default for F >= 256).  Warps claim small tiles of
//                  consecutive segments in order; a Y row is zero-filled inside the kernel by the tile
//                  that touches it first, a few microseconds before the first reduction reaches it, so
//                  the reductions hit dirty L2 lines and the row goes to DRAM once.  Ordering: publish
//                  flags per tile + per-32-tile counters, relaxed polls, bounded waits.
//   pull_kernel    gather-only TWO-PHASE form (default for F <= 128).  L2 reductions are several times
//                  slower than loads on this part (tools/replay.cu), so phase A writes the hyperedge
//                  features Xe and phase B gathers them per vertex through the CSR of H: plain stores
//                  only, every Y row written once, fixed summation order.
//   pc_kernel      CTA-level tiles with mbarrier producer / consumer warps (HGEF_PC=1; A/B only).
//
// All three stream the gathered rows through per-warp cp.async rings in shared memory with the row
// indices and per-row weights staged one tile ahead, so the only long-latency operations in the hot
// loops are the 128-bit feature-row copies themselves.  Heavy hyperedges (cut into w > 1 segments by
// the balancer) combine their partial sums with red.v4 into one row (scratch / Xe).
#include <cstdlib>
#include <type_traits>

#include "hgef_aggr.cuh"

namespace hg {
namespace {
using namespace dev;

constexpr uint32_t kFirst = 0x80000000u, kExcl = 0x40000000u, kIdMask = 0x3fffffffu;
constexpr int kIdxCap = 192;  // staged rows per warp tile (index, a_out, a_in); beyond that: global reads

struct FusedArgs {
  Args a;
  const int32_t *cflag, *iso;
  int32_t *ctrl;        // [0] tile counter, [2] give-up flag, [8 + b] #finished tiles of block b
                        // (32 tiles per block), [8 + nblk + t] done flag of tile t
  int64_t niso;
  int32_t ntiles, tile_segs, nblk;
  // TIMING EXPERIMENT ONLY (HGEF_UNSAFE_NOSYNC=1): skip publish / wait
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
__device__ __forceinline__ void st_relaxed(int *p, int v) {
  asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_inc(int *p) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
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

__device__ __forceinline__ void cp_async16(float *smem_dst, const float *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc) : "memory");
}
__device__ __forceinline__ void st_zero_v4(float *p) {
  asm volatile("st.global.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(0.0f) : "memory");
}
__device__ __forceinline__ void st_v4(float *p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// One SUB-WARP of SW lanes walks the rows of its run of segments; lane l of the sub-warp holds
// columns l*4 + j*SW*4 .. +3, j < VPL.  SW = 32 (whole warp per row, no divergence) for F >= 128.
template <int SW, int VPL>
struct Geo {
  static constexpr int kSub = 32 / SW;        // row streams per warp
  static constexpr int kColStride = SW * 4;   // floats between a lane's vectors
  static constexpr int kUnroll = (SW * VPL >= 64 ? 12 : (SW * VPL >= 32 ? 8 : 4)) / VPL;  // 16-byte vectors per lane and step:
                                                               // 12 (rows >= 1 KB), 8 (>= 512 B), else 4
};

constexpr int kTileMax = 32;   // segments per warp tile (one staging lane per segment)
constexpr int kHdrInts = 100;  // key[33] slot[32] scale[32], padded to 16 B
constexpr int kBufInts = kHdrInts + 3 * kIdxCap;  // + idx[] wout[] win[]

template <int SW, int VPL, bool EXACT>
__device__ __forceinline__ bool col_ok(int col, int j, int F) {
  return EXACT || col + j * Geo<SW, VPL>::kColStride < F;
}

// Persistent kernel, every WARP autonomous (no CTA-wide barrier anywhere).  A warp claims small
// tiles of consecutive segments from the global counter and runs a two-deep software pipeline:
//     claim t'  ->  stage t' (segment bounds, flagged column indices, per-row scales) in its
//     private shared-memory buffer  ->  zero-fill the rows t' touches first  ->  publish (one
//     release fence)  ->  wait until every tile <= t is published  ->  stream tile t (claimed
//     one iteration earlier): gather / reduce / scatter  ->  t = t'
// so the wait is normally already satisfied.  Inside a tile the X rows stream through a private
// shared-memory ring with cp.async, the loads of step k+1 issued before step k is consumed, and
// every per-row quantity the hot loops need (index, flags, a_in, a_out) comes from shared
// memory, so the only long-latency operations in flight are the feature rows themselves.
template <int SW, int VPL, bool EXACT>
__global__ void __launch_bounds__(kThreads) fused_kernel(const FusedArgs fa) {
  using G = Geo<SW, VPL>;
  extern __shared__ __align__(16) int32_t smem[];
  const Args &a = fa.a;
  const int T = fa.tile_segs;                      // <= 31
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t *wbuf = smem + warp * 2 * kBufInts;      // this warp's two tile buffers
  float *ring = reinterpret_cast<float *>(smem + kWarpsPerBlock * 2 * kBufInts) +
                warp * (2 * G::kUnroll * VPL * 32 * 4);  // this warp's cp.async landing ring
  const int sub = lane / SW;
  const int col = (lane % SW) * 4;
  const int F = a.F;
  const int64_t S = a.nwork;
  int *counter = fa.ctrl, *blk_cnt = fa.ctrl + 8, *done = fa.ctrl + 8 + fa.nblk;
  int blk_wm = 0;  // blocks [0, blk_wm) are known to be completely zero-filled

  auto claim = [&]() {
    int t = 0;
    if (lane == 0) t = atomicAdd(counter, 1);
    return __shfl_sync(kFull, t, 0);
  };
  // stage tile t into buffer b, zero-fill its first-touched rows, publish
  auto produce = [&](int t, int b) {
    int32_t *s_key = wbuf + b * kBufInts, *s_slot = s_key + 33;
    float *s_scale = reinterpret_cast<float *>(s_slot + 32);
    int32_t *s_idx = s_key + kHdrInts;
    float *s_wout = reinterpret_cast<float *>(s_idx + kIdxCap), *s_win = s_wout + kIdxCap;
    const int64_t s0 = (int64_t)t * T;
    const int nseg = (int)min((int64_t)T, S - s0);
    int32_t k = 0;
    if (lane <= nseg) k = __ldg(a.key + s0 + lane);
    if (lane < nseg) {
      s_slot[lane] = __ldg(a.seg_slot + s0 + lane);
      s_scale[lane] = edge_scale(a, __ldg(a.seg_edge + s0 + lane));
    }
    if (lane <= nseg) s_key[lane] = k;
    const int32_t p0 = __shfl_sync(kFull, k, 0), p1 = __shfl_sync(kFull, k, nseg);
    const int nidx = min(p1 - p0, kIdxCap);
    for (int i = lane; i < nidx; i += 32) {
      const uint32_t c = (uint32_t)__ldg(fa.cflag + p0 + i);
      s_idx[i] = (int32_t)c;
      s_wout[i] = a.a_out ? __ldg(a.a_out + (c & kIdMask)) : 1.0f;
      s_win[i] = a.a_in ? __ldg(a.a_in + (c & kIdMask)) : 1.0f;
    }
    __syncwarp();
    // (32 rows' flags are tested at once; only the flagged rows -- about a third -- cost a store pass)
    for (int32_t pb = p0; pb < p1; pb += 32) {
      const int32_t p = pb + lane;
      uint32_t c = 0;
      if (p < p1) c = (uint32_t)(p - p0 < kIdxCap ? s_idx[p - p0] : __ldg(fa.cflag + p));
      unsigned m = __ballot_sync(kFull, (c & kFirst) && !(c & kExcl));
      while (m) {  // kSub flagged rows per pass, one per sub-warp
        int b = -1;
#pragma unroll
        for (int g = 0; g < G::kSub; ++g) {
          if (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            if (g == sub) b = bit;
          }
        }
        const uint32_t cv = __shfl_sync(kFull, c, b < 0 ? 0 : b);
        if (b >= 0) {
          float *yp = a.Y + (int64_t)(cv & kIdMask) * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (col_ok<SW, VPL, EXACT>(col, j, F)) st_zero_v4(yp + j * G::kColStride);
        }
      }
    }
    // ... and this tile's share of the vertices that no hyperedge touches
    const int64_t i0 = fa.niso * t / fa.ntiles, i1 = fa.niso * (t + 1) / fa.ntiles;
    for (int64_t ib = i0; ib < i1; ib += 32) {
      const int n = (int)min((int64_t)32, i1 - ib);
      const int32_t my_v = lane < n ? __ldg(fa.iso + ib + lane) : 0;
      for (int r0 = 0; r0 < n; r0 += G::kSub) {
        const int r = r0 + sub;
        const int32_t v = __shfl_sync(kFull, my_v, r & 31);
        if (r < n) {
          float *yp = a.Y + (int64_t)v * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (col_ok<SW, VPL, EXACT>(col, j, F)) st_zero_v4(yp + j * G::kColStride);
        }
      }
    }
    __syncwarp();  // the lanes' zero stores are ordered before lane 0's release
    if (lane == 0) {
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      st_relaxed(done + t, 1);
      red_relaxed_inc(blk_cnt + (t >> 5));
    }
  };

  int t = claim(), cb = 0;
  if (t >= fa.ntiles) return;
  produce(t, 0);
  for (;;) {
    const int t_next = claim();
    if (t_next < fa.ntiles) produce(t_next, cb ^ 1);

    // every row tile t can touch was first-touched by a tile <= t: wait for all of them
    wait_zero_fill(fa.ctrl, fa.nblk, t, lane, blk_wm);
    const int32_t *s_key = wbuf + cb * kBufInts, *s_slot = s_key + 33;
    const float *s_scale = reinterpret_cast<const float *>(s_slot + 32);
    const int32_t *s_idx = s_key + kHdrInts;
    const float *s_wout = reinterpret_cast<const float *>(s_idx + kIdxCap), *s_win = s_wout + kIdxCap;
    const int nseg = (int)min((int64_t)T, S - (int64_t)t * T);
    const int32_t p0 = s_key[0];
    // per-row lookups: shared memory; global only for tiles that exceed the staging capacity
    const bool staged = s_key[nseg] - p0 <= kIdxCap;   // warp-uniform
    auto run_tile = [&](auto all_staged) {
    constexpr bool kStaged = decltype(all_staged)::value;
    auto row_idx = [&](int32_t p) -> uint32_t {
      return (uint32_t)((kStaged || p - p0 < kIdxCap) ? s_idx[p - p0] : __ldg(fa.cflag + p));
    };
    auto row_win = [&](int32_t p, uint32_t v) -> float {
      return (kStaged || p - p0 < kIdxCap) ? s_win[p - p0] : (a.a_in ? __ldg(a.a_in + v) : 1.0f);
    };
    auto row_wout = [&](int32_t p, uint32_t v) -> float {
      return (kStaged || p - p0 < kIdxCap) ? s_wout[p - p0] : (a.a_out ? __ldg(a.a_out + v) : 1.0f);
    };

    // Split the tile's segments into kSub contiguous runs of about equal row count, one per
    // sub-warp: the rows of a run are consecutive positions [plo, phi) of the index list.
    int i_lo = 0, i_hi = nseg;
    if (G::kSub > 1) {
      const int32_t k_l = lane < nseg ? s_key[lane] : 0x7fffffff;
      const int32_t rows = s_key[nseg] - p0;
#pragma unroll
      for (int g = 1; g < G::kSub; ++g) {
        const int b = __popc(__ballot_sync(kFull, k_l < p0 + (int32_t)(((int64_t)rows * g) / G::kSub)));
        if (g == sub) i_lo = b;
        if (g == sub + 1) i_hi = b;
      }
    }
    const int32_t plo = s_key[i_lo], phi = s_key[i_hi];
    auto ring_slot = [&](int buf, int u, int j) {
      return ring + ((((buf * G::kUnroll) + u) * VPL + j) * 32 + lane) * 4;
    };
    auto issue = [&](int buf, int32_t p) {
#pragma unroll
      for (int u = 0; u < G::kUnroll; ++u) {
        if (p + u < phi) {
          const float *xp = a.X + (int64_t)(row_idx(p + u) & kIdMask) * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (col_ok<SW, VPL, EXACT>(col, j, F)) cp_async16(ring_slot(buf, u, j), xp + j * G::kColStride);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int i = i_lo;
    while (i < i_hi && s_key[i + 1] == s_key[i]) ++i;  // (empty segments carry no work)
    int32_t seg_end = i < i_hi ? s_key[i + 1] : phi;
    Acc<VPL> acc;
    acc.zero();
    int buf = 0;
    issue(0, plo);
    for (int32_t p = plo; p < phi; p += G::kUnroll) {
      issue(buf ^ 1, p + G::kUnroll);              // (commits an empty group past the end)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll
      for (int u = 0; u < G::kUnroll; ++u) {
        if (p + u < phi) {
          const float w = a.a_in ? row_win(p + u, row_idx(p + u) & kIdMask) : 1.0f;
#pragma unroll
          for (int j = 0; j < VPL; ++j) {
            if (col_ok<SW, VPL, EXACT>(col, j, F)) {
              const float4 x = *reinterpret_cast<const float4 *>(ring_slot(buf, u, j));
              acc.v[j].x = fmaf(w, x.x, acc.v[j].x);
              acc.v[j].y = fmaf(w, x.y, acc.v[j].y);
              acc.v[j].z = fmaf(w, x.z, acc.v[j].z);
              acc.v[j].w = fmaf(w, x.w, acc.v[j].w);
            }
          }
          if (p + u + 1 == seg_end) {  // last member of segment i: finish it
            const int32_t slot = s_slot[i];
            if (slot >= 0) {           // heavy hyperedge: publish the partial sum, pass 2 scatters
              float *sp = a.scratch + (int64_t)slot * F + col;
#pragma unroll
              for (int j = 0; j < VPL; ++j)
                if (col_ok<SW, VPL, EXACT>(col, j, F)) red_add_v4(sp + j * G::kColStride, acc.v[j]);
            } else {
              scale_acc<VPL>(acc, s_scale[i]);
#pragma unroll 2
              for (int32_t q = s_key[i]; q < seg_end; ++q) {
                const uint32_t c = row_idx(q);
                const float o = row_wout(q, c & kIdMask);
                float *yp = a.Y + (int64_t)(c & kIdMask) * F + col;
                if (c & kExcl) {
#pragma unroll
                  for (int j = 0; j < VPL; ++j)
                    if (col_ok<SW, VPL, EXACT>(col, j, F))
                      st_v4(yp + j * G::kColStride, make_float4(acc.v[j].x * o, acc.v[j].y * o, acc.v[j].z * o,
                                                                acc.v[j].w * o));
                } else {
#pragma unroll
                  for (int j = 0; j < VPL; ++j)
                    if (col_ok<SW, VPL, EXACT>(col, j, F))
                      red_add_v4(yp + j * G::kColStride, make_float4(acc.v[j].x * o, acc.v[j].y * o,
                                                                     acc.v[j].z * o, acc.v[j].w * o));
                }
              }
            }
            acc.zero();
            ++i;
            while (i < i_hi && s_key[i + 1] == s_key[i]) ++i;
            seg_end = i < i_hi ? s_key[i + 1] : phi;
          }
        }
      }
      buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    };  // run_tile
    if (staged) run_tile(std::true_type{});
    else run_tile(std::false_type{});
    __syncwarp();  // all lanes are done with buffer cb before it is refilled
    if (t_next >= fa.ntiles) break;
    t = t_next;
    cb ^= 1;
  }
}

// ---------------------------------------------------------------------------------------
// Producer / consumer form.  The warp-autonomous kernel above keeps (#warps x 2) tiles in flight,
// hundreds of MB of rows between zero-fill and last reduction, so at F >= 128 about half of the Y
// rows are evicted and re-fetched (DRAM traffic 1.4-1.8x algorithmic, profiles/traffic.json), and
// every warp pays the serial claim -> stage -> fence -> publish chain itself.  Here a CTA owns a
// few LARGE tiles at a time:
//   PRODUCER warps (NP of NW) claim tiles in order, stage them in shared memory (segment bounds,
//     flagged indices, a_in / a_out per member), zero-fill the rows the tile touches first, fence,
//     publish, wait until every earlier tile is published, and hand the tile to the consumers
//     through an mbarrier.  They never have reductions in flight, so their fence is cheap.
//   CONSUMER warps take runs of segments of the current tile from a shared cursor and stream
//     them through their private cp.async ring exactly as above; they never touch global
//     synchronisation state and never fence.
// Tiles in flight: (#CTAs x NB) instead of (#warps x 2).
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
}

constexpr int kPcSegs = 128;   // max segments per CTA tile
constexpr int kPcRows = 640;   // staged rows per CTA tile (beyond that: global reads)
constexpr int kPcBufInts = (kPcSegs + 1 + 3) / 4 * 4 + 2 * kPcSegs + 3 * kPcRows;  // key | slot scale | idx wout win
constexpr int kPcRun = 4;      // segments a consumer warp takes from the cursor at a time (x row streams)

template <int SW, int VPL, bool EXACT, int NW, int NP, int NB>
__global__ void __launch_bounds__(NW * 32) pc_kernel(const FusedArgs fa) {
  using G = Geo<SW, VPL>;
  constexpr int NC = NW - NP;
  extern __shared__ __align__(16) int32_t smem[];
  __shared__ __align__(8) uint64_t s_full[NB], s_empty[NB];
  __shared__ int s_tile[NB], s_cursor[NB];
  const Args &a = fa.a;
  const int T = fa.tile_segs;                      // <= kPcSegs
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int sub = lane / SW;
  const int col = (lane % SW) * 4;
  const int F = EXACT ? SW * 4 * VPL : a.F;
  const int64_t S = a.nwork;
  if (tid == 0)
    for (int b = 0; b < NB; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_empty[b], NC); }
  __syncthreads();
  auto buf_key = [&](int b) { return smem + b * kPcBufInts; };

  if (warp < NP) {
    // ------------------------------ producer ------------------------------
    int *counter = fa.ctrl, *blk_cnt = fa.ctrl + 8, *done = fa.ctrl + 8 + fa.nblk;
    int blk_wm = 0;
    for (int it = warp;; it += NP) {
      const int b = it % NB;
      if (it >= NB) mbar_wait(&s_empty[b], ((it / NB) - 1) & 1);
      int t = 0;
      if (lane == 0) t = atomicAdd(counter, 1);
      t = __shfl_sync(kFull, t, 0);
      if (t >= fa.ntiles) {
        if (lane == 0) { s_tile[b] = -1; mbar_arrive(&s_full[b]); }
        break;
      }
      int32_t *s_key = buf_key(b), *s_slot = s_key + (kPcSegs + 1 + 3) / 4 * 4;
      float *s_scale = reinterpret_cast<float *>(s_slot + kPcSegs);
      int32_t *s_idx = reinterpret_cast<int32_t *>(s_scale + kPcSegs);
      float *s_wout = reinterpret_cast<float *>(s_idx + kPcRows), *s_win = s_wout + kPcRows;
      const int64_t s0 = (int64_t)t * T;
      const int nseg = (int)min((int64_t)T, S - s0);
      const int32_t p0 = __ldg(a.key + s0), p1 = __ldg(a.key + s0 + nseg);
      for (int i = lane; i <= nseg; i += 32) s_key[i] = __ldg(a.key + s0 + i);
      for (int i = lane; i < nseg; i += 32) {
        s_slot[i] = __ldg(a.seg_slot + s0 + i);
        s_scale[i] = edge_scale(a, __ldg(a.seg_edge + s0 + i));
      }
      const int nidx = min(p1 - p0, kPcRows);
      for (int i = lane; i < nidx; i += 32) {
        const uint32_t c = (uint32_t)__ldg(fa.cflag + p0 + i);
        s_idx[i] = (int32_t)c;
        s_wout[i] = a.a_out ? __ldg(a.a_out + (c & kIdMask)) : 1.0f;
        s_win[i] = a.a_in ? __ldg(a.a_in + (c & kIdMask)) : 1.0f;
      }
      __syncwarp();
      for (int32_t pb = p0; pb < p1; pb += 32) {        // zero-fill first-touched rows (flags 32 at a time)
        const int32_t p = pb + lane;
        uint32_t c = 0;
        if (p < p1) c = (uint32_t)(p - p0 < kPcRows ? s_idx[p - p0] : __ldg(fa.cflag + p));
        unsigned m = __ballot_sync(kFull, (c & kFirst) && !(c & kExcl));
        while (m) {
          int bsel = -1;
#pragma unroll
          for (int g = 0; g < G::kSub; ++g) {
            if (m) {
              const int bit = __ffs(m) - 1;
              m &= m - 1;
              if (g == sub) bsel = bit;
            }
          }
          const uint32_t cv = __shfl_sync(kFull, c, bsel < 0 ? 0 : bsel);
          if (bsel >= 0) {
            float *yp = a.Y + (int64_t)(cv & kIdMask) * F + col;
#pragma unroll
            for (int j = 0; j < VPL; ++j)
              if (col_ok<SW, VPL, EXACT>(col, j, F)) st_zero_v4(yp + j * G::kColStride);
          }
        }
      }
      const int64_t i0 = fa.niso * t / fa.ntiles, i1 = fa.niso * (t + 1) / fa.ntiles;
      for (int64_t ib = i0; ib < i1; ib += 32) {        // this tile's share of the isolated vertices
        const int n = (int)min((int64_t)32, i1 - ib);
        const int32_t my_v = lane < n ? __ldg(fa.iso + ib + lane) : 0;
        for (int r0 = 0; r0 < n; r0 += G::kSub) {
          const int r = r0 + sub;
          const int32_t v = __shfl_sync(kFull, my_v, r & 31);
          if (r < n) {
            float *yp = a.Y + (int64_t)v * F + col;
#pragma unroll
            for (int j = 0; j < VPL; ++j)
              if (col_ok<SW, VPL, EXACT>(col, j, F)) st_zero_v4(yp + j * G::kColStride);
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        st_relaxed(done + t, 1);
        red_relaxed_inc(blk_cnt + (t >> 5));
        s_tile[b] = t;
        s_cursor[b] = 0;
      }
      wait_zero_fill(fa.ctrl, fa.nblk, t, lane, blk_wm);   // every earlier tile is published too
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[b]);
    }
    return;
  }

  // ------------------------------ consumers ------------------------------
  float *ring = reinterpret_cast<float *>(smem + NB * kPcBufInts) + (warp - NP) * (2 * G::kUnroll * VPL * 32 * 4);
  bool finished[NP];
#pragma unroll
  for (int p = 0; p < NP; ++p) finished[p] = false;
  int nfinished = 0;
  for (int it = 0; nfinished < NP; ++it) {
    const int p = it % NP;
    bool skip = false;
#pragma unroll
    for (int q = 0; q < NP; ++q) if (q == p && finished[q]) skip = true;
    if (skip) continue;
    const int b = it % NB;
    mbar_wait(&s_full[b], (it / NB) & 1);
    const int t = s_tile[b];
    if (t < 0) {
#pragma unroll
      for (int q = 0; q < NP; ++q) if (q == p) finished[q] = true;
      ++nfinished;
      continue;
    }
    const int32_t *s_key = buf_key(b), *s_slot = s_key + (kPcSegs + 1 + 3) / 4 * 4;
    const float *s_scale = reinterpret_cast<const float *>(s_slot + kPcSegs);
    const int32_t *s_idx = reinterpret_cast<const int32_t *>(s_scale + kPcSegs);
    const float *s_wout = reinterpret_cast<const float *>(s_idx + kPcRows), *s_win = s_wout + kPcRows;
    const int nseg = (int)min((int64_t)T, S - (int64_t)t * T);
    const int32_t p0 = s_key[0];
    auto row_idx = [&](int32_t q) -> uint32_t {
      return (uint32_t)(q - p0 < kPcRows ? s_idx[q - p0] : __ldg(fa.cflag + q));
    };
    auto row_win = [&](int32_t q, uint32_t v) -> float {
      return q - p0 < kPcRows ? s_win[q - p0] : (a.a_in ? __ldg(a.a_in + v) : 1.0f);
    };
    auto row_wout = [&](int32_t q, uint32_t v) -> float {
      return q - p0 < kPcRows ? s_wout[q - p0] : (a.a_out ? __ldg(a.a_out + v) : 1.0f);
    };
    auto ring_slot = [&](int buf, int u, int j) {
      return ring + ((((buf * G::kUnroll) + u) * VPL + j) * 32 + lane) * 4;
    };
    for (;;) {
      int j0 = 0;
      if (lane == 0) j0 = atomicAdd(&s_cursor[b], kPcRun * G::kSub);
      j0 = __shfl_sync(kFull, j0, 0);
      if (j0 >= nseg) break;
      // this sub-warp's run of segments [i_lo, i_hi) and its rows [plo, phi)
      const int i_lo = min(j0 + sub * kPcRun, nseg), i_hi = min(i_lo + kPcRun, nseg);
      const int32_t plo = s_key[i_lo], phi = s_key[i_hi];
      auto issue = [&](int buf, int32_t q) {
#pragma unroll
        for (int u = 0; u < G::kUnroll; ++u) {
          if (q + u < phi) {
            const float *xp = a.X + (int64_t)(row_idx(q + u) & kIdMask) * F + col;
#pragma unroll
            for (int j = 0; j < VPL; ++j)
              if (col_ok<SW, VPL, EXACT>(col, j, F)) cp_async16(ring_slot(buf, u, j), xp + j * G::kColStride);
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      int i = i_lo;
      while (i < i_hi && s_key[i + 1] == s_key[i]) ++i;
      int32_t seg_end = i < i_hi ? s_key[i + 1] : phi;
      Acc<VPL> acc;
      acc.zero();
      int buf = 0;
      issue(0, plo);
      for (int32_t q = plo; q < phi; q += G::kUnroll) {
        issue(buf ^ 1, q + G::kUnroll);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll
        for (int u = 0; u < G::kUnroll; ++u) {
          if (q + u < phi) {
            const float w = a.a_in ? row_win(q + u, row_idx(q + u) & kIdMask) : 1.0f;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
              if (col_ok<SW, VPL, EXACT>(col, j, F)) {
                const float4 x = *reinterpret_cast<const float4 *>(ring_slot(buf, u, j));
                acc.v[j].x = fmaf(w, x.x, acc.v[j].x);
                acc.v[j].y = fmaf(w, x.y, acc.v[j].y);
                acc.v[j].z = fmaf(w, x.z, acc.v[j].z);
                acc.v[j].w = fmaf(w, x.w, acc.v[j].w);
              }
            }
            if (q + u + 1 == seg_end) {
              const int32_t slot = s_slot[i];
              if (slot >= 0) {
                float *sp = a.scratch + (int64_t)slot * F + col;
#pragma unroll
                for (int j = 0; j < VPL; ++j)
                  if (col_ok<SW, VPL, EXACT>(col, j, F)) red_add_v4(sp + j * G::kColStride, acc.v[j]);
              } else {
                scale_acc<VPL>(acc, s_scale[i]);
#pragma unroll 2
                for (int32_t r = s_key[i]; r < seg_end; ++r) {
                  const uint32_t c = row_idx(r);
                  const float o = row_wout(r, c & kIdMask);
                  float *yp = a.Y + (int64_t)(c & kIdMask) * F + col;
                  if (c & kExcl) {
#pragma unroll
                    for (int j = 0; j < VPL; ++j)
                      if (col_ok<SW, VPL, EXACT>(col, j, F))
                        st_v4(yp + j * G::kColStride, make_float4(acc.v[j].x * o, acc.v[j].y * o, acc.v[j].z * o,
                                                                  acc.v[j].w * o));
                  } else {
#pragma unroll
                    for (int j = 0; j < VPL; ++j)
                      if (col_ok<SW, VPL, EXACT>(col, j, F))
                        red_add_v4(yp + j * G::kColStride, make_float4(acc.v[j].x * o, acc.v[j].y * o,
                                                                       acc.v[j].z * o, acc.v[j].w * o));
                  }
                }
              }
              acc.zero();
              ++i;
              while (i < i_hi && s_key[i + 1] == s_key[i]) ++i;
              seg_end = i < i_hi ? s_key[i + 1] : phi;
            }
          }
        }
        buf ^= 1;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[b]);
  }
}

// ---------------------------------------------------------------------------------------
// PULL form: no reductions at all.  Replaying the address stream of the workload with a trivially
// lean kernel (tools/replay.cu, profiles/r01_replay_address_stream.txt) shows where the time of every
// scatter-based variant goes: gathering the 2.2 M member rows takes 85 us, issuing the same number
// of red.v4 takes 310-410 us.  Loads are several times cheaper than L2 reductions on this part, so
// the aggregation is done as two gather passes over the same streaming machinery:
//   phase A  (units = balancer segments)   Xe[e] = s1[e] s2[e] * sum_{u in seg} a_in[u] X[u]
//            light hyperedge: one plain 128-bit store per 16 B; heavy (w > 1): red into the
//            pre-zeroed Xe row (rare)
//   phase B  (units = vertices, CSR of H)  Y[v]  = a_out[v] * sum_{e in H[v]} Xe[e]
//            plain stores, every Y row written exactly once, no zero-fill, no ordering protocol,
//            and a fixed summation order per row (bit-reproducible run to run).
// Xe ([M, F]) goes through L2 / HBM between the phases: M < N rows, and a replica's worth of it is
// L2-resident when phase B reads it.  DRAM bytes ~ 8FN + 8FM instead of 8FN, but at gather speed.
struct PullArgs {
  const int32_t *ptr;       // unit offsets: A = balancer key [S+1], B = H indptr [N+1]
  const int32_t *ind;       // gathered row of every position: A = H^T colind (vertex), B = H colind (hyperedge)
  const int32_t *out_row;   // A: seg_edge [S]; B: NULL (the unit is the row)
  const int32_t *slot;      // A: seg_slot [S] (>= 0: heavy, reduce into the pre-zeroed row); B: NULL
  const float *src;         // A: X; B: Xe
  const float *w_in;        // A: a_in (or NULL); B: NULL
  const float *w_out1, *w_out2;   // A: s1, s2 per hyperedge; B: a_out per vertex, NULL
  float *dst;               // A: Xe; B: Y
  int32_t *counter;
  int64_t nunit;
  int32_t F, tile;
};

constexpr int kPullUnits = 96;   // units (segments / vertices) per warp tile
constexpr int kPullHdr = (kPullUnits + 1 + 2 * kPullUnits + 3) / 4 * 4;   // key[T+1] orow[T] oscale[T]
constexpr int kPullBuf = kPullHdr + 2 * kIdxCap;                          // + idx[] win[]

template <int SW, int VPL>
struct PullGeo {   // ring depth: 8 (rows >= 512 B) or 4 sixteen-byte vectors per lane and step
  static constexpr int kUnroll = (SW * VPL >= 32 ? 8 : 4) / VPL;
};

template <int SW, int VPL, bool EXACT>
__global__ void __launch_bounds__(kThreads) pull_kernel(const PullArgs pa) {
  using G = Geo<SW, VPL>;
  constexpr int U = PullGeo<SW, VPL>::kUnroll;
  extern __shared__ __align__(16) int32_t smem[];
  const int T = pa.tile;                                    // <= kPullUnits
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t *wbuf = smem + warp * 2 * kPullBuf;
  float *ring = reinterpret_cast<float *>(smem + kWarpsPerBlock * 2 * kPullBuf) + warp * (2 * U * VPL * 32 * 4);
  const int sub = lane / SW;
  const int col = (lane % SW) * 4;
  const int F = EXACT ? SW * 4 * VPL : pa.F;
  const int64_t ntile = (pa.nunit + T - 1) / T;

  auto claim = [&]() {
    int t = 0;
    if (lane == 0) t = atomicAdd(pa.counter, 1);
    return __shfl_sync(kFull, t, 0);
  };
  // stage tile t into buffer b: unit bounds, output row / scale per unit, gathered-row indices and weights
  auto stage = [&](int t, int b) {
    int32_t *s_key = wbuf + b * kPullBuf, *s_orow = s_key + kPullUnits + 1;
    float *s_oscale = reinterpret_cast<float *>(s_orow + kPullUnits);
    int32_t *s_idx = s_key + kPullHdr;
    float *s_win = reinterpret_cast<float *>(s_idx + kIdxCap);
    const int64_t u0 = (int64_t)t * T;
    const int nu = (int)min((int64_t)T, pa.nunit - u0);
    const int32_t p0 = __ldg(pa.ptr + u0), p1 = __ldg(pa.ptr + u0 + nu);
    for (int i = lane; i < nu; i += 32) {
      const int32_t k = __ldg(pa.ptr + u0 + i), k_next = __ldg(pa.ptr + u0 + i + 1);
      const int32_t orow = pa.out_row ? __ldg(pa.out_row + u0 + i) : (int32_t)(u0 + i);
      float sc = pa.w_out1 ? __ldg(pa.w_out1 + orow) : 1.0f;
      if (pa.w_out2) sc *= __ldg(pa.w_out2 + orow);
      const bool heavy = pa.slot && __ldg(pa.slot + u0 + i) >= 0;
      s_key[i] = k;
      s_orow[i] = heavy ? ~orow : orow;                     // (negative = heavy: reduce, do not store)
      s_oscale[i] = sc;
      // a unit with no member produces a zero row (phase B: vertices in no hyperedge)
      if (k_next == k && !pa.out_row)
        for (int c0 = 0; c0 < F; c0 += 4) st_zero_v4(pa.dst + (int64_t)orow * F + c0);
    }
    if (lane == 0) s_key[nu] = p1;
    const int nidx = min(p1 - p0, kIdxCap);
    for (int i = lane; i < nidx; i += 32) {
      const int32_t v = __ldg(pa.ind + p0 + i);
      s_idx[i] = v;
      s_win[i] = pa.w_in ? __ldg(pa.w_in + v) : 1.0f;
    }
    __syncwarp();
  };

  int t = claim(), cb = 0;
  if (t >= ntile) return;
  stage(t, 0);
  for (;;) {
    const int t_next = claim();
    if (t_next < ntile) stage(t_next, cb ^ 1);
    const int32_t *s_key = wbuf + cb * kPullBuf, *s_orow = s_key + kPullUnits + 1;
    const float *s_oscale = reinterpret_cast<const float *>(s_orow + kPullUnits);
    const int32_t *s_idx = s_key + kPullHdr;
    const float *s_win = reinterpret_cast<const float *>(s_idx + kIdxCap);
    const int nu = (int)min((int64_t)T, pa.nunit - (int64_t)t * T);
    const int32_t p0 = s_key[0];
    auto row_idx = [&](int32_t q) -> int32_t { return q - p0 < kIdxCap ? s_idx[q - p0] : __ldg(pa.ind + q); };
    auto row_win = [&](int32_t q, int32_t v) -> float {
      return q - p0 < kIdxCap ? s_win[q - p0] : (pa.w_in ? __ldg(pa.w_in + v) : 1.0f);
    };
    // contiguous runs of units per sub-warp, balanced by row count (first unit whose start >= target)
    auto first_unit_at = [&](int32_t target) {
      int lo = 0, hi = nu;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_key[mid] < target) lo = mid + 1; else hi = mid;
      }
      return lo;
    };
    const int32_t rows = s_key[nu] - p0;
    const int i_lo = G::kSub > 1 && sub > 0 ? first_unit_at(p0 + (int32_t)(((int64_t)rows * sub) / G::kSub)) : 0;
    const int i_hi = G::kSub > 1 && sub + 1 < G::kSub
                         ? first_unit_at(p0 + (int32_t)(((int64_t)rows * (sub + 1)) / G::kSub)) : nu;
    const int32_t plo = s_key[i_lo], phi = s_key[i_hi];
    auto ring_slot = [&](int buf, int u, int j) { return ring + ((((buf * U) + u) * VPL + j) * 32 + lane) * 4; };
    auto issue = [&](int buf, int32_t q) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (q + u < phi) {
          const float *xp = pa.src + (int64_t)row_idx(q + u) * F + col;
#pragma unroll
          for (int j = 0; j < VPL; ++j)
            if (col_ok<SW, VPL, EXACT>(col, j, F)) cp_async16(ring_slot(buf, u, j), xp + j * G::kColStride);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int i = i_lo;
    while (i < i_hi && s_key[i + 1] == s_key[i]) ++i;
    int32_t unit_end = i < i_hi ? s_key[i + 1] : phi;
    Acc<VPL> acc;
    acc.zero();
    int buf = 0;
    issue(0, plo);
    for (int32_t q = plo; q < phi; q += U) {
      issue(buf ^ 1, q + U);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (q + u < phi) {
          const float w = pa.w_in ? row_win(q + u, row_idx(q + u)) : 1.0f;
#pragma unroll
          for (int j = 0; j < VPL; ++j) {
            if (col_ok<SW, VPL, EXACT>(col, j, F)) {
              const float4 x = *reinterpret_cast<const float4 *>(ring_slot(buf, u, j));
              acc.v[j].x = fmaf(w, x.x, acc.v[j].x);
              acc.v[j].y = fmaf(w, x.y, acc.v[j].y);
              acc.v[j].z = fmaf(w, x.z, acc.v[j].z);
              acc.v[j].w = fmaf(w, x.w, acc.v[j].w);
            }
          }
          if (q + u + 1 == unit_end) {       // unit i complete: one output row
            const int32_t orow = s_orow[i];
            scale_acc<VPL>(acc, s_oscale[i]);
            float *op = pa.dst + (int64_t)(orow < 0 ? ~orow : orow) * F + col;
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
              if (col_ok<SW, VPL, EXACT>(col, j, F)) {
                if (orow < 0) red_add_v4(op + j * G::kColStride, acc.v[j]);
                else st_v4(op + j * G::kColStride, acc.v[j]);
              }
            }
            acc.zero();
            ++i;
            while (i < i_hi && s_key[i + 1] == s_key[i]) ++i;
            unit_end = i < i_hi ? s_key[i + 1] : phi;
          }
        }
      }
      buf ^= 1;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    if (t_next >= ntile) break;
    t = t_next;
    cb ^= 1;
  }
}

// zero the Xe rows of heavy hyperedges (their segments reduce into them in phase A)
__global__ void pull_zero_heavy_kernel(int64_t nheavy_segs, const int32_t *__restrict__ heavy_segs,
                                       const int32_t *__restrict__ seg_edge, float *__restrict__ xe, int F) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (w >= nheavy_segs) return;
  float *row = xe + (int64_t)seg_edge[heavy_segs[w]] * F;
  for (int c = lane; c < F; c += 32) row[c] = 0.0f;
}

template <int SW, int VPL, bool EXACT>
int launch_pull_phase(const hgPlan *plan, PullArgs &pa, cudaStream_t s) {
  using G = Geo<SW, VPL>;
  const double avg_len = (double)plan->nnz / (double)pa.nunit;
  int T = (int)(0.75 * kIdxCap / (avg_len > 1.0 ? avg_len : 1.0));
  if (T > kPullUnits) T = kPullUnits;
  if (T < G::kSub) T = G::kSub;
  pa.tile = T;
  const size_t smem = (size_t)kWarpsPerBlock * 2 * kPullBuf * sizeof(int32_t) +
                      (size_t)kWarpsPerBlock * 2 * PullGeo<SW, VPL>::kUnroll * VPL * 32 * sizeof(float4);
  auto kern = pull_kernel<SW, VPL, EXACT>;
  HG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  static const int ctas_env = getenv("HGEF_PULL_CTAS") ? atoi(getenv("HGEF_PULL_CTAS")) : 0;
  if (ctas_env > 0 && ctas_env < per_sm) per_sm = ctas_env;
  int64_t grid = (int64_t)plan->sm_count * per_sm;
  const int64_t max_useful = ceil_div<int64_t>(ceil_div<int64_t>(pa.nunit, T), kWarpsPerBlock);
  if (grid > max_useful) grid = max_useful;
  kern<<<(unsigned)grid, kThreads, smem, s>>>(pa);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

// Pass 2 for heavy hyperedges with the single-writer flags (rows that were never zero-filled
// must be stored, not reduced).
template <int SW, int VPL>
__global__ void __launch_bounds__(kThreads) fused_pass2_kernel(const FusedArgs fa) {
  using G = Geo<SW, VPL>;
  const Args &a = fa.a;
  const int lane = threadIdx.x & 31;
  const int sub = lane / SW;
  const int col = (lane % SW) * 4;
  const int F = a.F;
  const int64_t nsub = (int64_t)gridDim.x * kWarpsPerBlock * G::kSub;
  for (int64_t i = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * G::kSub + sub; i < a.nwork;
       i += nsub) {
    const int32_t s = __ldg(a.seg_list + i);
    const int32_t lo = __ldg(a.key + s), hi = __ldg(a.key + s + 1);
    const float *sp = a.scratch + (int64_t)__ldg(a.seg_slot + s) * F + col;
    const float sc = edge_scale(a, __ldg(a.seg_edge + s));
    Acc<VPL> acc;
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
      acc.v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col + j * G::kColStride < F) acc.v[j] = *reinterpret_cast<const float4 *>(sp + j * G::kColStride);
    }
    scale_acc<VPL>(acc, sc);
#pragma unroll 4
    for (int32_t p = lo; p < hi; ++p) {
      const uint32_t c = (uint32_t)__ldg(fa.cflag + p);
      const uint32_t v = c & kIdMask;
      const float o = a.a_out ? __ldg(a.a_out + v) : 1.0f;
      float *yp = a.Y + (int64_t)v * F + col;
#pragma unroll
      for (int j = 0; j < VPL; ++j) {
        if (col + j * G::kColStride < F) {
          const float4 val = make_float4(acc.v[j].x * o, acc.v[j].y * o, acc.v[j].z * o, acc.v[j].w * o);
          if (c & kExcl) st_v4(yp + j * G::kColStride, val);
          else red_add_v4(yp + j * G::kColStride, val);
        }
      }
    }
  }
}

template <int SW, int VPL, bool EXACT>
int launch(hgPlan *plan, FusedArgs &fa, cudaStream_t s) {
  using G = Geo<SW, VPL>;
  const int F = fa.a.F;
  // warp-tile size: ~tile_kb of gathered + scattered rows, a multiple of the row streams of a warp,
  // at most one staging lane per segment (31) and, on average, within the staging capacity
  const double avg_len = (double)plan->nnz / (double)plan->nseg;
  static const double tile_env = getenv("HGEF_TILE_KB") ? atof(getenv("HGEF_TILE_KB")) : 0.0;
  const double tile_kb = tile_env > 0 ? tile_env : (F >= 128 ? 128.0 : 64.0);
  int T = (int)(tile_kb * 1024.0 / (avg_len * 4.0 * F * 2.0));
  const int cap_T = (int)(0.75 * kIdxCap / avg_len);
  if (T > cap_T) T = cap_T;
  T = T / G::kSub * G::kSub;
  if (T > 31) T = 31 / G::kSub * G::kSub;
  if (T < G::kSub) T = G::kSub;
  fa.tile_segs = T;
  fa.ntiles = (int32_t)ceil_div<int64_t>(plan->nseg, T);
  fa.nblk = (fa.ntiles + 31) / 32;
  const size_t smem = (size_t)kWarpsPerBlock * 2 * kBufInts * sizeof(int32_t) +
                      (size_t)kWarpsPerBlock * 2 * G::kUnroll * VPL * 32 * sizeof(float4);
  HG_CUDA_TRY(cudaFuncSetAttribute(fused_kernel<SW, VPL, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fused_kernel<SW, VPL, EXACT>, kThreads, smem));
  if (per_sm < 1) per_sm = 1;
  static const int ctas_env = getenv("HGEF_CTAS_PER_SM") ? atoi(getenv("HGEF_CTAS_PER_SM")) : 0;
  // Wide rows (F >= 256) carry enough bytes per warp; fewer resident warps keep the set of
  // zero-filled-but-not-yet-reduced rows inside L2 (measured: DRAM traffic 2.0x -> ~1.5x algorithmic)
  const int ctas_cap = ctas_env > 0 ? ctas_env : (F >= 256 ? 1 : 0);
  if (ctas_cap > 0 && ctas_cap < per_sm) per_sm = ctas_cap;
  int64_t grid = (int64_t)plan->sm_count * per_sm;
  const int64_t max_useful = ceil_div<int64_t>(fa.ntiles, kWarpsPerBlock);
  if (grid > max_useful) grid = max_useful;
  HG_CUDA_TRY(cudaMemsetAsync(plan->ctrl, 0, (size_t)(fa.ntiles + fa.nblk + 8) * sizeof(int32_t), s));
  fa.ctrl = plan->ctrl;
  fa.a.nwork = plan->nseg;
  fused_kernel<SW, VPL, EXACT><<<(unsigned)grid, kThreads, smem, s>>>(fa);
  HG_CUDA_TRY(cudaGetLastError());
  if (plan->nheavy_segs > 0) {
    fa.a.nwork = plan->nheavy_segs;
    fa.a.seg_list = plan->heavy_segs;
    int64_t g2 = ceil_div<int64_t>(plan->nheavy_segs, kWarpsPerBlock * G::kSub);
    if (g2 > (int64_t)plan->sm_count * 8) g2 = (int64_t)plan->sm_count * 8;
    fused_pass2_kernel<SW, VPL><<<(unsigned)g2, kThreads, 0, s>>>(fa);
    HG_CUDA_TRY(cudaGetLastError());
  }
  return HG_OK;
}

template <int SW, int VPL, bool EXACT, int NW, int NP, int NB>
int launch_pc(hgPlan *plan, FusedArgs &fa, cudaStream_t s) {
  using G = Geo<SW, VPL>;
  const int F = fa.a.F;
  const double avg_len = (double)plan->nnz / (double)plan->nseg;
  static const double tile_env = getenv("HGEF_PC_TILE_KB") ? atof(getenv("HGEF_PC_TILE_KB")) : 0.0;
  const double tile_kb = tile_env > 0 ? tile_env : 256.0;
  int T = (int)(tile_kb * 1024.0 / (avg_len * 4.0 * F * 2.0));
  const int cap_T = (int)(0.8 * kPcRows / avg_len);
  if (T > cap_T) T = cap_T;
  if (T > kPcSegs) T = kPcSegs;
  if (T < 8) T = 8;
  fa.tile_segs = T;
  fa.ntiles = (int32_t)ceil_div<int64_t>(plan->nseg, T);
  fa.nblk = (fa.ntiles + 31) / 32;
  const size_t smem = (size_t)NB * kPcBufInts * sizeof(int32_t) +
                      (size_t)(NW - NP) * 2 * G::kUnroll * VPL * 32 * sizeof(float4);
  auto kern = pc_kernel<SW, VPL, EXACT, NW, NP, NB>;
  HG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NW * 32, smem));
  HG_REQUIRE(per_sm >= 1, "launch_pc: kernel does not fit an SM (smem %zu)", smem);
  static const int ctas_env = getenv("HGEF_CTAS_PER_SM") ? atoi(getenv("HGEF_CTAS_PER_SM")) : 0;
  if (ctas_env > 0 && ctas_env < per_sm) per_sm = ctas_env;
  int64_t grid = (int64_t)plan->sm_count * per_sm;
  if (grid > fa.ntiles) grid = fa.ntiles;
  HG_CUDA_TRY(cudaMemsetAsync(plan->ctrl, 0, (size_t)(fa.ntiles + fa.nblk + 8) * sizeof(int32_t), s));
  fa.ctrl = plan->ctrl;
  fa.a.nwork = plan->nseg;
  kern<<<(unsigned)grid, NW * 32, smem, s>>>(fa);
  HG_CUDA_TRY(cudaGetLastError());
  if (plan->nheavy_segs > 0) {
    fa.a.nwork = plan->nheavy_segs;
    fa.a.seg_list = plan->heavy_segs;
    int64_t g2 = ceil_div<int64_t>(plan->nheavy_segs, kWarpsPerBlock * G::kSub);
    if (g2 > (int64_t)plan->sm_count * 8) g2 = (int64_t)plan->sm_count * 8;
    fused_pass2_kernel<SW, VPL><<<(unsigned)g2, kThreads, 0, s>>>(fa);
    HG_CUDA_TRY(cudaGetLastError());
  }
  return HG_OK;
}

template <int SW, int VPL, bool EXACT>
int launch_pull(hgPlan *plan, const dev::Args &a, cudaStream_t s) {
  const int F = a.F;
  HG_CUDA_TRY(cudaMemsetAsync(plan->ctrl, 0, 8 * sizeof(int32_t), s));
  if (plan->nheavy_segs > 0) {
    pull_zero_heavy_kernel<<<(unsigned)ceil_div<int64_t>(plan->nheavy_segs * 32, 256), 256, 0, s>>>(
        plan->nheavy_segs, plan->heavy_segs, plan->seg_edge, plan->xe, F);
    HG_CUDA_TRY(cudaGetLastError());
  }
  PullArgs A{};
  A.ptr = plan->key; A.ind = plan->colind; A.out_row = plan->seg_edge; A.slot = plan->seg_slot;
  A.src = a.X; A.w_in = a.a_in; A.w_out1 = a.s1; A.w_out2 = a.s2; A.dst = plan->xe;
  A.counter = plan->ctrl; A.nunit = plan->nseg; A.F = F;
  static const int only = getenv("HGEF_PULL_ONLY") ? atoi(getenv("HGEF_PULL_ONLY")) : 0;  // timing: 1 = A, 2 = B
  if (only != 2)
    if (int rc = launch_pull_phase<SW, VPL, EXACT>(plan, A, s)) return rc;
  if (only == 1) return HG_OK;
  PullArgs B{};
  B.ptr = plan->h_ptr; B.ind = plan->h_ind; B.src = plan->xe; B.w_out1 = a.a_out; B.dst = a.Y;
  B.counter = plan->ctrl + 4; B.nunit = plan->num_nodes; B.F = F;
  return launch_pull_phase<SW, VPL, EXACT>(plan, B, s);
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

// When to use the persistent form (measured, profiles/r01_tuning_sweeps.txt):
//  * Y must not fit the L2: below ~64 MB the memset and the reductions of the two-pass form never
//    leave the L2 and its one-warp-per-segment grid has the lower latency (literal Pubmed, F=128:
//    18 us vs 43 us; the crossover is at N*F*4 ~ 64 MB);
//  * segments must be short: the per-warp staging holds kIdxCap rows; with ngs-long segments (the
//    Walmart-shaped graph) the two-pass kernel's 32-wide index batches fit better (1.6 vs 3.1 ms).
bool fused_available(const hgPlan *plan, int F, bool force) {
  if (plan->cflag == nullptr || plan->ctrl == nullptr) return false;
  if (force) return true;
  if ((double)plan->num_nodes * F * 4.0 < 64.0 * 1048576.0) return false;
  return (double)plan->nnz / (double)plan->nseg <= 0.125 * kIdxCap;
}

// The pull form needs H (built by the plan), short units on both sides (staging) and no extreme
// vertex degree (one sub-warp walks a vertex's hyperedges).
bool pull_available(const hgPlan *plan, int F, bool force) {
  if (plan->h_ptr == nullptr || plan->ctrl == nullptr) return false;
  if (plan->max_vdeg > 4096) return false;
  if (force) return true;
  if ((double)plan->num_nodes * F * 4.0 < 64.0 * 1048576.0) return false;
  // measured crossover (profiles/r01_tuning_sweeps.txt): the gather-only form wins for rows up to 512 B
  // (F=32/64/128: 160/216/400 us vs 184/274/454 us), the scatter form from F=256 up (747 vs 754 us, 1410 vs 1683)
  if (F > 128) return false;
  return (double)plan->nnz / (double)plan->nseg <= 0.125 * kIdxCap;
}

int ensure_xe(hgPlan *plan, int F, cudaStream_t s);

int launch_pull_any(hgPlan *plan, const dev::Args &base, cudaStream_t s) {
  if (int rc = ensure_xe(plan, base.F, s)) return rc;
  const int F = base.F;
  int sw = F <= 16 ? 4 : (F <= 64 ? 8 : (F <= 128 ? 16 : 32));
  static const int sw_env = getenv("HGEF_PULL_SW") ? atoi(getenv("HGEF_PULL_SW")) : 0;
  if (sw_env == 4 || sw_env == 8 || sw_env == 16 || sw_env == 32) sw = sw_env;
  while (sw < 32 && F > sw * 16) sw *= 2;
  const int vpl = F <= sw * 4 ? 1 : (F <= sw * 8 ? 2 : 4);
  const bool exact = F == sw * 4 * vpl;
#define HG_CASE(SW_, VPL_)                                               \
  if (sw == SW_ && vpl == VPL_)                                          \
    return exact ? launch_pull<SW_, VPL_, true>(plan, base, s) : launch_pull<SW_, VPL_, false>(plan, base, s)
  HG_CASE(4, 1); HG_CASE(4, 2); HG_CASE(4, 4);
  HG_CASE(8, 1); HG_CASE(8, 2); HG_CASE(8, 4);
  HG_CASE(16, 1); HG_CASE(16, 2); HG_CASE(16, 4);
  HG_CASE(32, 1); HG_CASE(32, 2); HG_CASE(32, 4);
#undef HG_CASE
  return set_error(HG_EINVAL, "launch_pull: no kernel for F=%d", F);
}

// Y is NOT zero-filled by the caller; scratch (if any) is.
int launch_fused(hgPlan *plan, const dev::Args &base, cudaStream_t s) {
  FusedArgs fa{};
  fa.a = base;
  fa.cflag = plan->cflag;
  fa.iso = plan->iso_list;
  fa.niso = plan->niso;
  const int F = base.F;
  // sub-warp width: a whole warp per row from F = 128 up (no divergence between row streams),
  // narrower sub-warps below so that all 32 lanes still carry data
  static const int sw_env = getenv("HGEF_SW") ? atoi(getenv("HGEF_SW")) : 0;
  int sw = sw_env;
  if (sw != 4 && sw != 8 && sw != 16 && sw != 32) sw = F <= 16 ? 4 : (F <= 64 ? 8 : (F <= 128 ? 16 : 32));
  while (sw < 32 && F > sw * 16) sw *= 2;  // at most 4 vectors per lane
  const int vpl = F <= sw * 4 ? 1 : (F <= sw * 8 ? 2 : 4);
  const bool exact = F == sw * 4 * vpl;
  static const int pc = getenv("HGEF_PC") ? atoi(getenv("HGEF_PC")) : 0;   // A/B: 1 = 16 warps / 2 producers, 2 = 16 / 4
  if (pc && exact) {
#define HG_PC_CASE(SW_, VPL_)                                                                  \
    if (sw == SW_ && vpl == VPL_)                                                              \
      return pc == 2 ? launch_pc<SW_, VPL_, true, 16, 4, 8>(plan, fa, s)                        \
                     : launch_pc<SW_, VPL_, true, 16, 2, 4>(plan, fa, s)
    HG_PC_CASE(8, 1); HG_PC_CASE(8, 2); HG_PC_CASE(16, 2); HG_PC_CASE(32, 1); HG_PC_CASE(32, 2); HG_PC_CASE(32, 4);
#undef HG_PC_CASE
  }
#define HG_CASE(SW_, VPL_)                                               \
  if (sw == SW_ && vpl == VPL_)                                          \
    return exact ? launch<SW_, VPL_, true>(plan, fa, s) : launch<SW_, VPL_, false>(plan, fa, s)
  HG_CASE(4, 1); HG_CASE(4, 2); HG_CASE(4, 4);
  HG_CASE(8, 1); HG_CASE(8, 2); HG_CASE(8, 4);
  HG_CASE(16, 1); HG_CASE(16, 2); HG_CASE(16, 4);
  HG_CASE(32, 1); HG_CASE(32, 2); HG_CASE(32, 4);
#undef HG_CASE
  return set_error(HG_EINVAL, "launch_fused: no kernel for F=%d", F);
}

}  // namespace hg
