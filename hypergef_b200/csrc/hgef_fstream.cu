// hgef_fstream.cu -- the FUSED STREAM form: both stages of the aggregation in ONE persistent launch as
// register-only row streams, the hyperedge features handed from stage A to stage B through the L2 and
// discarded there before they are written back.
//
// Why: the two-launch stream form (hgef_stream.cu) runs both stages at 76-96 % of the copy peak in DRAM terms
// and still tops out at ~70 % of the algorithmic roofline, because `Xe` makes a round trip through HBM (1.3x
// the algorithmic traffic).  profiles/r02_l2probe_dram.txt: a written line survives 65 MB (plain) to 130 MB
// (with eviction-priority hints) of streaming traffic in the B200 L2, and `discard.global.L2` after its last
// read removes the write-back: DRAM bytes = algorithmic bytes.  profiles/r02_tma_probe.txt: TMA bulk row
// copies are served ~1 per 300 clocks per issuing warp, so rows below 1 KB are gathered with LDG here (the
// ring form, hgef_ring.cu, moves rows >= 1 KB with cp.async.bulk).
// The reference keeps the hyperedge feature in a register of the thread that scatters it
// (hgnnaggr_cuda.cu:26-45) and pays with scalar atomics into Y; here every Y row is written once, plainly.
//
// Structure.  The row programs of the stream form (src / dst words per position) are cut into ITEMS of a few
// KB of rows; DISCARD items list the hyperedges whose last stage-B reader lies in one block of B items.  One
// ticket order (hgef_ring.cu: fused_get_sched): a B item follows the A items that produce its hyperedge
// features (+ lag), a discard item follows the B items that read its rows (+ lag).  Every warp is an
// autonomous worker: it claims tickets (the next ticket and its item record are fetched while the current
// item streams), waits for an item's dependencies (relaxed polls of per-block completion counters, then ONE
// acquire fence), streams the item with the inner loop of the stream form (sub-warp row streams, index words
// one per lane, two alternating register sets of row vectors), and publishes a finished item with
// fence.release.gpu + a relaxed increment (no L1 invalidate on the release side).
// Loads: X read-only with an L2 evict-first hint; Xe plain (written earlier in this launch: coherent after
// the acquire fence, never rewritten within a launch).  Stores: Xe evict-last, Y evict-first.
// Deadlock freedom: tickets are claimed in order by running warps only; an item waits only for items with
// smaller tickets; A items wait for nothing.  Waits are bounded (give-up flag -> hg_plan_check).
#include "hgef_stream.cuh"

namespace hg {
namespace {
using namespace dev;

enum { kPolNormal = 0, kPolFirst = 1, kPolLast = 2 };

struct FArgs {
  const int32_t *src[2], *dst[2];   // row programs: [0] stage A, [1] stage B
  const float *in[2];
  float *out[2];
  const float *w_in;                // gather-side weight of stage A (a_in) or null
  const float *w_o1[2], *w_o2[2];   // output scales per output row
  const int4 *items;                // two words per ticket (hgef_plan.cuh)
  const int32_t *dperm;             // hyperedges in discard order
  const int32_t *iso;               // vertices in no hyperedge: Y row = 0
  int32_t *ctrl;
  int32_t niso, nitem, nslab, slabF, F;
  int32_t nblkA, nblkB, GA, GB;
  int32_t track_b;                  // B items are counted too (discard items wait for them)
  int32_t pol_x, pol_xe_w, pol_y;
  int32_t batch;                    // tickets claimed per atomic; finished items are published once per batch
  int32_t debug;                    // bit 0: no release fence, bit 1: no dependency waits (timing experiments only)
  const int4 *tabs;                 // split-role form: item records per kind and index
  int32_t GC, maxlead, doff;        // discard items; how far (in items) stage A may run ahead of the claimed B items;
                                    // how many B blocks later a block's discards are attempted
};

__device__ __forceinline__ int ld_relaxed(const int *p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_inc_relaxed(int *p) {
  asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ uint64_t make_policy(int kind) {
  uint64_t p;
  if (kind == kPolFirst) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == kPolLast) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// X rows: read-only path with an L2 eviction hint.  Xe rows: written earlier in THIS launch -- a plain
// (coherent) load, volatile so that it stays behind the acquire fence of the dependency wait.
template <int STAGE>
__device__ __forceinline__ float4 ld_row16(const char *p, uint64_t pol) {
  float4 v;
  if (STAGE == 0)
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  else
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_row_hint(float *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}

// KV = 128-bit row loads in flight per lane: 8 at three CTAs per SM, up to 32 at one ("fat" warps: the same
// bytes in flight per SM from a third of the warps, so that the set of items the grid works on stays small)
template <int SW, int VPL, int KV>
struct SGeo {
  static constexpr int kSub = 32 / SW;                               // row streams per warp
  static constexpr int kU = (KV / VPL) < SW ? (KV / VPL) : SW;       // rows in flight per stream
  static constexpr int kHB = kU >= 2 ? kU / 2 : 1;                   // rows per half-batch
  static constexpr int kStride = SW * 4;                             // floats between a lane's vectors
};

// One sub-warp of SW lanes streams positions [ps, pe) of stage STAGE (a whole number of units): index words
// of SW positions one per lane (the next chunk's are prefetched), rows loaded in half-batches into two
// alternating register sets, one output row per unit end.
template <int STAGE, int SW, int VPL, int KV, bool HAS_WIN, bool PIPE>
__device__ __forceinline__ void stream_run(const FArgs &fa, int32_t ps, int32_t pe, int col0, int Fs, int sl,
                                           uint64_t pol_in, uint64_t pol_out) {
  using G = SGeo<SW, VPL, KV>;
  constexpr int HB = PIPE ? G::kHB : G::kU, NH = SW / HB;
  static_assert(!PIPE || (NH >= 2 && NH % 2 == 0), "a chunk is a whole number of half-batch pairs");
  const int col = sl * 4;
  const uint32_t row_bytes = (uint32_t)fa.F * 4u;
  const float *in = fa.in[STAGE] + col0;
  float *out = fa.out[STAGE] + col0;
  const int32_t *__restrict__ src = fa.src[STAGE];
  const int32_t *__restrict__ dst = fa.dst[STAGE];
  const float *__restrict__ w_o1 = fa.w_o1[STAGE];
  const float *__restrict__ w_o2 = fa.w_o2[STAGE];
  constexpr bool WIN = HAS_WIN && STAGE == 0;
  uint32_t pat = 0;   // bit (stream * SW) for every row stream of the warp
#pragma unroll
  for (int q = 0; q < G::kSub; ++q) pat |= 1u << (q * SW);
  // column mask of this lane's vectors; a masked vector LOADS column 0 of the slab instead and is never stored
  bool ok[VPL];
  int off[VPL];
  const char *in_v[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) {
    ok[v] = col + v * G::kStride < Fs;
    off[v] = ok[v] ? col + v * G::kStride : 0;
    in_v[v] = reinterpret_cast<const char *>(in + off[v]);
  }
  float4 acc[VPL];
#pragma unroll
  for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  int32_t cb = ps;
  uint32_t c_src = 0, c_dst = 0;
  auto fetch = [&](int32_t p, uint32_t &ws, uint32_t &wdst) {
    ws = (uint32_t)__ldg(src + p);
    wdst = (uint32_t)__ldg(dst + p);
  };
  if (cb + sl < pe) fetch(cb + sl, c_src, c_dst);
  float4 x0[HB][VPL], x1[HB][VPL];
  // (positions past the end of the run carry src word 0: row 0 is loaded and never used for output -- a run
  //  ends with a unit end, which resets `acc`)
  auto load_half = [&](float4(&x)[HB][VPL], uint32_t words, int j0) {
#pragma unroll
    for (int u = 0; u < HB; ++u) {
      const uint32_t id = __shfl_sync(kFull, words, j0 + u, SW);
      const uint64_t rb = (uint64_t)id * row_bytes;
#pragma unroll
      for (int v = 0; v < VPL; ++v) x[u][v] = ld_row16<STAGE>(in_v[v] + rb, pol_in);
    }
  };
  if (PIPE) load_half(x0, c_src, 0);
  while (__any_sync(kFull, cb < pe)) {
    float c_w = 1.0f, c_sc = 1.0f;
    if (cb + sl < pe) {
      if (WIN) c_w = __ldg(fa.w_in + c_src);
      if (c_dst & kEnd) {
        const uint32_t orow = c_dst & kRowMask;
        if (w_o1) c_sc = __ldg(w_o1 + orow);
        if (w_o2) c_sc *= __ldg(w_o2 + orow);
      }
    }
    const uint32_t endm = __ballot_sync(kFull, (c_dst & kEnd) != 0);
    uint32_t n_src = 0, n_dst = 0;   // next chunk's words: in flight while this chunk streams
    if (cb + SW + sl < pe) fetch(cb + SW + sl, n_src, n_dst);
    auto consume_half = [&](const float4(&x)[HB][VPL], int j0) {
#pragma unroll
      for (int u = 0; u < HB; ++u) {
        const int j = j0 + u;
        float w = 1.0f;
        if (WIN) w = __shfl_sync(kFull, c_w, j, SW);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
          if (WIN) {
            acc[v].x = fmaf(w, x[u][v].x, acc[v].x);
            acc[v].y = fmaf(w, x[u][v].y, acc[v].y);
            acc[v].z = fmaf(w, x[u][v].z, acc[v].z);
            acc[v].w = fmaf(w, x[u][v].w, acc[v].w);
          } else {
            acc[v].x += x[u][v].x;
            acc[v].y += x[u][v].y;
            acc[v].z += x[u][v].z;
            acc[v].w += x[u][v].w;
          }
        }
        if (endm & (pat << j)) {   // warp-uniform: some stream finishes a unit at this position
          const uint32_t d = __shfl_sync(kFull, c_dst, j, SW);
          const float sc = __shfl_sync(kFull, c_sc, j, SW);
          if (d & kEnd) {          // this stream does: one output row
            char *op = reinterpret_cast<char *>(out) + (uint64_t)(d & kIdMask) * row_bytes;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
              if (ok[v]) {
                const float4 r = make_float4(acc[v].x * sc, acc[v].y * sc, acc[v].z * sc, acc[v].w * sc);
                float *o = reinterpret_cast<float *>(op) + off[v];
                if (STAGE == 0 && (d & kHeavy)) red_add_v4(o, r);
                else st_row_hint(o, r, pol_out);
              }
              acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
        }
      }
    };
    if constexpr (PIPE) {
#pragma unroll 1
      for (int h = 0; h < NH; h += 2) {
        if (h > 0 && !__any_sync(kFull, cb + h * HB < pe)) break;
        load_half(x1, c_src, (h + 1) * HB);
        consume_half(x0, h * HB);
        if (h + 2 < NH) load_half(x0, c_src, (h + 2) * HB);
        else load_half(x0, n_src, 0);                      // first half of the next chunk
        consume_half(x1, (h + 1) * HB);
      }
    } else {
#pragma unroll 1
      for (int h = 0; h < NH; ++h) {
        if (h > 0 && !__any_sync(kFull, cb + h * HB < pe)) break;
        load_half(x0, c_src, h * HB);
        consume_half(x0, h * HB);
      }
    }
    cb += SW;
    c_src = n_src;
    c_dst = n_dst;
  }
}

template <int SW, int VPL, int KV, bool HAS_WIN, int MINB, bool PIPE>
__global__ void __launch_bounds__(kThreads, MINB) fstream_kernel(const __grid_constant__ FArgs fa) {
  using G = SGeo<SW, VPL, KV>;
  static_assert(G::kSub <= 4, "an item record carries three split points");
  const int lane = threadIdx.x & 31;
  const int sub = lane / SW, sl = lane % SW;
  const int total = fa.nitem * fa.nslab;
  const uint64_t pol_x = make_policy(fa.pol_x), pol_xe_w = make_policy(fa.pol_xe_w), pol_y = make_policy(fa.pol_y);
  int wm0 = 0, wm1 = 0, wm_slab = -1;   // completion watermarks: A blocks (B items need them), B blocks (discards)
  bool gave_up = false;

  // tickets are claimed `batch` at a time: the single counter serves ~300 M same-address atomics a second, which
  // is less than one per item at wide rows (result valid in lane 0 only; broadcast where it is used)
  const int batch = fa.batch;
  auto claim = [&]() -> int {
    int t = 0;
    if (lane == 0) t = atomicAdd(fa.ctrl, batch);
    return t;
  };
  // Finished items [pend_lo, pend_hi) are published together: ONE release fence (it drains the stores of the
  // whole warp), then one relaxed increment per item.  Always before this warp starts to wait for anything, so a
  // waiting warp never sits on a finished, unpublished item (that keeps the wait graph acyclic).
  int pend_lo = 0, pend_hi = 0;
  auto publish = [&]() {
    if (pend_hi <= pend_lo) return;
    __syncwarp();
    const int tt = pend_lo + lane;
    if (tt < pend_hi) {
      const int sl_ = fa.nslab > 1 ? tt / fa.nitem : 0;
      const int4 rec = __ldg(fa.items + 2 * (tt - sl_ * fa.nitem));
      const int kd = rec.z & 3, ix = rec.z >> 2;
      if (kd == kKindA || (kd == kKindB && fa.track_b)) {
        if (!(fa.debug & 1)) asm volatile("fence.release.gpu;" ::: "memory");
        red_inc_relaxed(fa.ctrl + kCntOff + sl_ * (fa.nblkA + fa.nblkB) + (kd == kKindA ? 0 : fa.nblkA) + ix / kBlk);
      }
    }
    __syncwarp();
    pend_lo = pend_hi;
  };
  // all blocks [0, need) of one kind complete?  relaxed polls, one acquire fence at the end
  auto wait_blocks = [&](int kind, int slab, int need) {
    if (slab != wm_slab) { wm_slab = slab; wm0 = wm1 = 0; }
    int w = kind == 0 ? wm0 : wm1;
    if (w >= need || (fa.debug & 2)) return;
    publish();
    const int G_ = kind == 0 ? fa.GA : fa.GB;
    const int *cnt = fa.ctrl + kCntOff + (int64_t)slab * (fa.nblkA + fa.nblkB) + (kind == 0 ? 0 : fa.nblkA);
    unsigned spins = 0;
    while (w < need && !gave_up) {
      const int b = w + lane;
      bool done = true;
      if (b < need) done = ld_relaxed(cnt + b) == min(kBlk, G_ - b * kBlk);
      const unsigned m = __ballot_sync(kFull, done);
      w = min(need, w + (m == kFull ? 32 : __ffs(~m) - 1));
      if (w < need && m != kFull) {
        __nanosleep(64);
        ++spins;
        // bounded: a protocol bug must not hang the GPU; once one item gave up, nobody waits any more
        if (spins > (1u << 20) || ((spins & 255u) == 0 && ld_relaxed(fa.ctrl + kFlagOff) != 0)) {
          if (lane == 0) atomicExch(fa.ctrl + kFlagOff, 1);
          gave_up = true;
        }
      }
    }
    if (kind == 0) wm0 = w; else wm1 = w;
    asm volatile("fence.acquire.gpu;" ::: "memory");
  };

  int t = __shfl_sync(kFull, claim(), 0);
  if (t >= total) return;
  int t_end = t + batch;                 // end of the batch this ticket belongs to
  pend_lo = pend_hi = t;
  int slab = fa.nslab > 1 ? t / fa.nitem : 0;
  int4 ia = __ldg(fa.items + 2 * (t - slab * fa.nitem)), ib = make_int4(0, 0, 0, 0);
  if (G::kSub > 1) ib = __ldg(fa.items + 2 * (t - slab * fa.nitem) + 1);
  int nb_raw = claim();                  // the next batch: claimed one batch ahead
  for (;;) {
    // the next ticket and its record: in flight while this item streams
    int tn = t + 1;
    if (tn == t_end) {
      tn = __shfl_sync(kFull, nb_raw, 0);
      t_end = tn + batch;
      nb_raw = claim();
    }
    int4 na = make_int4(0, 0, 0, 0), nb = na;
    int nslab_i = 0;
    if (tn < total) {
      nslab_i = fa.nslab > 1 ? tn / fa.nitem : 0;
      na = __ldg(fa.items + 2 * (tn - nslab_i * fa.nitem));
      if (G::kSub > 1) nb = __ldg(fa.items + 2 * (tn - nslab_i * fa.nitem) + 1);
    }

    const int kind = ia.z & 3, idx = ia.z >> 2;
    const int col0 = slab * fa.slabF;
    const int Fs = min(fa.slabF, fa.F - col0);
    const int blk_base = slab * (fa.nblkA + fa.nblkB);
    if (kind == kKindC) {
      // drop the consumed hyperedge rows from the L2 (no write-back); their last readers are complete
      wait_blocks(1, slab, ia.w);
      const int lines = Fs >> 5;                           // 128-byte lines per row slab
      const int totl = (ia.y - ia.x) * lines;
      for (int x = lane; x < totl; x += 32) {
        const int r = x / lines, l = x - r * lines;
        const int32_t e = __ldg(fa.dperm + ia.x + r);
        const float *p = fa.in[1] + (int64_t)e * fa.F + col0 + l * 32;
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
      }
    } else {
      // this sub-stream's share of the item (split points are unit-aligned run boundaries)
      int32_t ps = ia.x, pe = ia.y;
      if (G::kSub == 2) { ps = sub == 0 ? ia.x : ib.x; pe = sub == 0 ? ib.x : ia.y; }
      if (G::kSub == 4) {
        ps = sub == 0 ? ia.x : (sub == 1 ? ib.x : (sub == 2 ? ib.y : ib.z));
        pe = sub == 0 ? ib.x : (sub == 1 ? ib.y : (sub == 2 ? ib.z : ia.y));
      }
      if (kind == kKindB) {
        if (ia.w > 0) wait_blocks(0, slab, ia.w);
        // this item's share of the vertices that no hyperedge touches
        if (fa.niso > 0) {
          const int i0 = (int)((int64_t)fa.niso * idx / fa.GB), i1 = (int)((int64_t)fa.niso * (idx + 1) / fa.GB);
          for (int i = i0 + sub; i < i1; i += G::kSub) {
            float *yp = fa.out[1] + (int64_t)__ldg(fa.iso + i) * fa.F + col0;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              if (sl * 4 + v * G::kStride < Fs) st_row_hint(yp + sl * 4 + v * G::kStride, make_float4(0.f, 0.f, 0.f, 0.f), pol_y);
          }
        }
        stream_run<1, SW, VPL, KV, HAS_WIN, PIPE>(fa, ps, pe, col0, Fs, sl, 0, pol_y);
      } else {
        stream_run<0, SW, VPL, KV, HAS_WIN, PIPE>(fa, ps, pe, col0, Fs, sl, pol_x, pol_xe_w);
      }
    }
    pend_hi = t + 1;                        // this item is finished (discard items publish nothing)
    if (t + 1 == t_end || tn >= total || tn != t + 1) publish();
    if (tn >= total) break;
    if (tn != t + 1) pend_lo = pend_hi = tn;
    t = tn; slab = nslab_i; ia = na; ib = nb;
  }
}



// ---------------------------------------------------------------------------------------------------
// SPLIT-ROLE form.  The merged ticket order above makes every warp wait for whatever its next ticket needs; a
// lag that avoids the waits is longer than the L2 retains.  Here the CTAs of one launch take FIXED roles (even
// blocks: stage A, odd blocks: stage B + discards), each role with its own ticket counter, and the distance
// between the two is bounded from both sides: a B item waits until the A blocks it needs are complete (it
// cannot run ahead), an A item waits while it is more than `maxlead` items ahead of the completed B prefix (it
// cannot run away).  maxlead >= the largest lead any B item needs (computed from the graph, RingSched::lead), so
// the A items the oldest unfinished B item needs always pass the throttle: no deadlock.  No warp ever claims an
// item of the other role, so there is no contended "is it ready?" race.
// ---------------------------------------------------------------------------------------------------
template <int SW, int VPL, bool HAS_WIN, int MINB, bool PIPE>
__global__ void __launch_bounds__(kThreads, MINB) dstream_kernel(const __grid_constant__ FArgs fa) {
  using G = SGeo<SW, VPL, 8>;
  static_assert(G::kSub <= 4, "an item record carries three split points");
  const int lane = threadIdx.x & 31;
  const int sub = lane / SW, sl = lane % SW;
  const bool roleB = (blockIdx.x & 1) != 0;
  const uint64_t pol_x = make_policy(fa.pol_x), pol_xe_w = make_policy(fa.pol_xe_w), pol_y = make_policy(fa.pol_y);
  const int GA = fa.GA, GB = fa.GB, GC = fa.GC;
  const int nblk = fa.nblkA + fa.nblkB;
  int wmA = 0, wmB = 0, slabA = -1, slabB = -1;   // completed prefix blocks of the slab last looked at
  bool gave_up = false;

  // blocks [0, need) of `kind` of `slab` complete?  waits (bounded) until they are; one acquire fence at the end
  auto wait_blocks = [&](int kind, int slab, int need) {
    int &w = kind == 0 ? wmA : wmB;
    int &ws = kind == 0 ? slabA : slabB;
    if (slab != ws) { ws = slab; w = 0; }
    if (w >= need || (fa.debug & 2)) return;
    const int G_ = kind == 0 ? GA : GB;
    const int *cnt = fa.ctrl + kCntOff + (int64_t)slab * nblk + (kind == 0 ? 0 : fa.nblkA);
    unsigned spins = 0;
    while (w < need && !gave_up) {
      const int b = w + lane;
      bool done = true;
      if (b < need) done = ld_relaxed(cnt + b) == min(kBlk, G_ - b * kBlk);
      const unsigned m = __ballot_sync(kFull, done);
      w = min(need, w + (m == kFull ? 32 : __ffs(~m) - 1));
      if (w < need && m != kFull) {
        __nanosleep(100);
        ++spins;
        if (spins > (1u << 20) || ((spins & 255u) == 0 && ld_relaxed(fa.ctrl + kFlagOff) != 0)) {
          if (lane == 0) atomicExch(fa.ctrl + kFlagOff, 1);
          gave_up = true;
        }
      }
    }
    asm volatile("fence.acquire.gpu;" ::: "memory");
  };
  auto publish = [&](int kind, int slab, int idx) {
    __syncwarp();
    if (lane == 0) {
      if (!(fa.debug & 1)) asm volatile("fence.release.gpu;" ::: "memory");
      red_inc_relaxed(fa.ctrl + kCntOff + slab * nblk + (kind == 0 ? 0 : fa.nblkA) + idx / kBlk);
    }
  };
  auto bounds = [&](const int4 &ia, const int4 &ib, int32_t &ps, int32_t &pe) {
    ps = ia.x; pe = ia.y;
    if (G::kSub == 2) { ps = sub == 0 ? ia.x : ib.x; pe = sub == 0 ? ib.x : ia.y; }
    if (G::kSub == 4) {
      ps = sub == 0 ? ia.x : (sub == 1 ? ib.x : (sub == 2 ? ib.y : ib.z));
      pe = sub == 0 ? ib.x : (sub == 1 ? ib.y : (sub == 2 ? ib.z : ia.y));
    }
  };

  if (!roleB) {
    // ------------------------------ stage A warps ------------------------------
    const int total = GA * fa.nslab;
    int64_t seenB = 0;                    // B tickets seen claimed
    int t_raw = 0;
    if (lane == 0) t_raw = atomicAdd(fa.ctrl, 1);
    for (;;) {
      const int t = __shfl_sync(kFull, t_raw, 0);
      if (t >= total) break;
      if (lane == 0) t_raw = atomicAdd(fa.ctrl, 1);          // the next ticket: in flight while this item streams
      const int slab = fa.nslab > 1 ? t / GA : 0;
      const int idx = t - slab * GA;
      const int4 ia = __ldg(fa.tabs + 2 * idx);
      int4 ib = make_int4(0, 0, 0, 0);
      if (G::kSub > 1) ib = __ldg(fa.tabs + 2 * idx + 1);
      // throttle: not more than maxlead items ahead of the B items CLAIMED so far (the B ticket counter; relaxed
      // polls).  B items are claimed in order and a claimed B item only ever waits for A items below its own index
      // + the graph's lead <= maxlead, which pass this test: no deadlock.
      if (!(fa.debug & 2)) {
        const int64_t tb_need = (int64_t)slab * GB + idx - fa.maxlead;
        if (tb_need > seenB) {
          unsigned spins = 0;
          for (;;) {
            int v = 0;
            if (lane == 0) v = ld_relaxed(fa.ctrl + kNextB);
            seenB = __shfl_sync(kFull, v, 0);
            if (seenB >= tb_need || gave_up) break;
            __nanosleep(200);
            if (++spins > (1u << 20) || ((spins & 255u) == 0 && ld_relaxed(fa.ctrl + kFlagOff) != 0)) {
              if (lane == 0) atomicExch(fa.ctrl + kFlagOff, 1);
              gave_up = true;
            }
          }
        }
      }
      const int col0 = slab * fa.slabF;
      const int Fs = min(fa.slabF, fa.F - col0);
      int32_t ps, pe;
      bounds(ia, ib, ps, pe);
      stream_run<0, SW, VPL, 8, HAS_WIN, PIPE>(fa, ps, pe, col0, Fs, sl, pol_x, pol_xe_w);
      publish(0, slab, idx);
    }
  } else {
    // ------------------------------ stage B warps (and the discards) ------------------------------
    const int total = GB * fa.nslab;
    const int4 *tabB = fa.tabs + 2 * GA, *tabC = fa.tabs + 2 * (GA + GB);
    int t_raw = 0;
    if (lane == 0) t_raw = atomicAdd(fa.ctrl + kNextB, 1);
    for (;;) {
      const int t = __shfl_sync(kFull, t_raw, 0);
      if (t >= total) break;
      if (lane == 0) t_raw = atomicAdd(fa.ctrl + kNextB, 1);
      const int slab = fa.nslab > 1 ? t / GB : 0;
      const int idx = t - slab * GB;
      const int4 ia = __ldg(tabB + 2 * idx);
      int4 ib = make_int4(0, 0, 0, 0);
      if (G::kSub > 1) ib = __ldg(tabB + 2 * idx + 1);
      if (ia.w > 0) wait_blocks(0, slab, ia.w);
      const int col0 = slab * fa.slabF;
      const int Fs = min(fa.slabF, fa.F - col0);
      if (fa.niso > 0) {   // this item's share of the vertices that no hyperedge touches
        const int i0 = (int)((int64_t)fa.niso * idx / GB), i1 = (int)((int64_t)fa.niso * (idx + 1) / GB);
        for (int i = i0 + sub; i < i1; i += G::kSub) {
          float *yp = fa.out[1] + (int64_t)__ldg(fa.iso + i) * fa.F + col0;
#pragma unroll
          for (int v = 0; v < VPL; ++v)
            if (sl * 4 + v * G::kStride < Fs) st_row_hint(yp + sl * 4 + v * G::kStride, make_float4(0.f, 0.f, 0.f, 0.f), pol_y);
        }
      }
      int32_t ps, pe;
      bounds(ia, ib, ps, pe);
      stream_run<1, SW, VPL, 8, HAS_WIN, PIPE>(fa, ps, pe, col0, Fs, sl, 0, pol_y);
      if (fa.track_b) publish(1, slab, idx);
      // discards: the warp that finishes the LAST item of B block k looks after discard item k - doff (the rows whose
      // last reader lies in that earlier block) -- exactly one attempt per discard item, no contended claim.  The
      // attempt is non-blocking: if the B blocks up to it are not all complete yet, the item is skipped (a discard is an
      // optimisation: a skipped one costs a write-back, not correctness).  A warp waiting here would sit on its
      // prefetched B ticket, which the block it waits for may contain.
      if (GC > 0 && (idx % kBlk == kBlk - 1 || idx == GB - 1)) {
        const int ci = idx / kBlk - fa.doff;
        if (ci >= 0) {
          if (slab != slabB) { slabB = slab; wmB = 0; }
          const int *cnt = fa.ctrl + kCntOff + (int64_t)slab * nblk + fa.nblkA;
          for (int pass = 0; pass < 64 && wmB < ci + 1; ++pass) {     // polling passes over the missing blocks
            const int bb = wmB + lane;
            bool done = true;
            if (bb < ci + 1) done = ld_relaxed(cnt + bb) == min(kBlk, GB - bb * kBlk);
            const unsigned m = __ballot_sync(kFull, done);
            const int adv = m == kFull ? 32 : __ffs(~m) - 1;
            wmB = min(ci + 1, wmB + adv);
            if (adv < 32) break;
          }
          if (wmB >= ci + 1) {
            asm volatile("fence.acquire.gpu;" ::: "memory");
            const int4 ic = __ldg(tabC + 2 * ci);
            const int lines = Fs >> 5;
            const int totl = (ic.y - ic.x) * lines;
            for (int x = lane; x < totl; x += 32) {
              const int r = x / lines, l = x - r * lines;
              const int32_t e = __ldg(fa.dperm + ic.x + r);
              const float *p = fa.in[1] + (int64_t)e * fa.F + col0 + l * 32;
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
            }
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// ALTERNATING form.  Every warp takes an A item, then a B item, then an A item ... from the two ticket counters of the
// split-role tables.  A items never wait.  A warp HOLDS its B ticket until the A blocks the item needs are complete
// and meanwhile keeps taking A items (it never polls while there is A work), so stage A runs ahead of stage B by
// exactly what the graph needs plus the items in flight, with no throttle and no lag parameter.  Only when the A
// tickets are exhausted does a warp wait for its B item -- and then every A item is in the hands of a warp that is
// streaming it, so the wait ends: no deadlock, and not every CTA has to be resident.
// Measured (profiles/r02_alternating.txt): what decides whether Xe stays in the L2 is the size of the set of claimed,
// unfinished items = warps x item bytes.  8 KB items: DRAM bytes 0.97x algorithmic (Xe never reaches HBM, X rows
// are re-used in the L2); 16 KB: 1.26x; 32 KB: 1.36x.  But an item is an isolated load pipeline (fill + drain = two
// memory latencies) plus ~400 control instructions, so at 8 KB the launch takes 908 us against 586 us for the
// two-launch form; fetching the item record / index words ahead (cp.async into a per-warp shared-memory slot) and
// publishing late were built and bought 2 % -- and publishing late puts Xe back into HBM (one more round of items
// in flight).  The form that ships is still the two-launch one.
// ---------------------------------------------------------------------------------------------------
template <int SW, int VPL, int KV, bool HAS_WIN, int MINB, bool PIPE>
__global__ void __launch_bounds__(kThreads, MINB) astream_kernel(const __grid_constant__ FArgs fa) {
  using G = SGeo<SW, VPL, KV>;
  static_assert(G::kSub <= 4, "an item record carries three split points");
  const int lane = threadIdx.x & 31;
  const int sub = lane / SW, sl = lane % SW;
  const uint64_t pol_x = make_policy(fa.pol_x), pol_xe_w = make_policy(fa.pol_xe_w), pol_y = make_policy(fa.pol_y);
  const int GA = fa.GA, GB = fa.GB, GC = fa.GC;
  const int nblk = fa.nblkA + fa.nblkB;
  const int totalA = GA * fa.nslab, totalB = GB * fa.nslab;
  const int4 *tabB = fa.tabs + 2 * GA, *tabC = fa.tabs + 2 * (GA + GB);
  int wmA = 0, wmB = 0, slabA = -1, slabB = -1;   // completed prefix blocks of the slab last looked at
  bool gave_up = false;

  // How far is the prefix of complete blocks of `kind`?  (non-blocking)  The prefix found is shared through one
  // word per (slab, kind): a warp starts its scan where the furthest scan of any warp ended, so a check is one
  // load in the common case instead of a walk over everything finished since this warp last looked (that walk
  // cost 20-25 % of the launch: profiles/r02_alternating.txt, section 1 vs 2).
  auto scan_blocks = [&](int kind, int slab, int need) -> bool {
    int &w = kind == 0 ? wmA : wmB;
    int &ws = kind == 0 ? slabA : slabB;
    if (slab != ws) { ws = slab; w = 0; }
    if ((fa.debug & 2) || w >= need) return true;
    const int G_ = kind == 0 ? GA : GB;
    const int nb = kind == 0 ? fa.nblkA : fa.nblkB;
    int *gw = fa.ctrl + kCntOff + (int64_t)nblk * fa.nslab + slab * 64 + kind * 32;
    const int g = ld_relaxed(gw);
    if (g > w) w = g;
    if (w >= need) return true;
    const int *cnt = fa.ctrl + kCntOff + (int64_t)slab * nblk + (kind == 0 ? 0 : fa.nblkA);
    // (the scan runs past `need` to the end of the complete prefix: the next item's need is usually a little further)
    for (int pass = 0; pass < 8; ++pass) {
      const int b = w + lane;
      bool done = false;
      if (b < nb) done = ld_relaxed(cnt + b) == min(kBlk, G_ - b * kBlk);
      const unsigned m = __ballot_sync(kFull, done);
      const int adv = m == kFull ? 32 : __ffs(~m) - 1;
      w += adv;
      if (adv < 32) break;
    }
    if (w > g && lane == 0) {
      __threadfence();
      atomicMax(gw, w);
    }
    return w >= need;
  };
  auto publish = [&](int kind, int slab, int idx) {
    __syncwarp();
    if (lane == 0) {
      if (!(fa.debug & 1)) asm volatile("fence.release.gpu;" ::: "memory");
      red_inc_relaxed(fa.ctrl + kCntOff + slab * nblk + (kind == 0 ? 0 : fa.nblkA) + idx / kBlk);
    }
  };
  auto bounds = [&](const int4 &ia, const int4 &ib, int32_t &ps, int32_t &pe) {
    ps = ia.x; pe = ia.y;
    if (G::kSub == 2) { ps = sub == 0 ? ia.x : ib.x; pe = sub == 0 ? ib.x : ia.y; }
    if (G::kSub == 4) {
      ps = sub == 0 ? ia.x : (sub == 1 ? ib.x : (sub == 2 ? ib.y : ib.z));
      pe = sub == 0 ? ib.x : (sub == 1 ? ib.y : (sub == 2 ? ib.z : ia.y));
    }
  };

  int ta_raw = 0, tb_raw = 0;
  if (lane == 0) {
    ta_raw = atomicAdd(fa.ctrl, 1);
    tb_raw = atomicAdd(fa.ctrl + kNextB, 1);
  }
  int tb = -1;            // the B ticket this warp holds (-1: its claim is in flight in tb_raw)
  bool b_claimed = true;  // tb_raw holds a claim that has not been looked at yet
  bool b_none = false;    // the B tickets are exhausted
  int4 ba = make_int4(0, 0, 0, 0), bb = ba;
  for (;;) {
    // ------------------------------ an A item (never waits) ------------------------------
    const int ta = __shfl_sync(kFull, ta_raw, 0);
    const bool didA = ta < totalA;
    if (didA) {
      if (lane == 0) ta_raw = atomicAdd(fa.ctrl, 1);          // the next ticket: in flight while this item streams
      const int slab = fa.nslab > 1 ? ta / GA : 0;
      const int idx = ta - slab * GA;
      const int4 ia = __ldg(fa.tabs + 2 * idx);
      int4 ib = make_int4(0, 0, 0, 0);
      if (G::kSub > 1) ib = __ldg(fa.tabs + 2 * idx + 1);
      const int col0 = slab * fa.slabF;
      const int Fs = min(fa.slabF, fa.F - col0);
      int32_t ps, pe;
      bounds(ia, ib, ps, pe);
      stream_run<0, SW, VPL, KV, HAS_WIN, PIPE>(fa, ps, pe, col0, Fs, sl, pol_x, pol_xe_w);
      publish(0, slab, idx);
    }
    // ------------------------------ the B item this warp holds, if it is ready ------------------------------
    if (b_claimed) {
      b_claimed = false;
      tb = __shfl_sync(kFull, tb_raw, 0);
      if (tb >= totalB) { tb = -1; b_none = true; }
      else {
        const int idx = tb - (fa.nslab > 1 ? tb / GB : 0) * GB;
        ba = __ldg(tabB + 2 * idx);
        if (G::kSub > 1) bb = __ldg(tabB + 2 * idx + 1);
      }
    }
    if (tb >= 0) {
      const int slab = fa.nslab > 1 ? tb / GB : 0;
      const int idx = tb - slab * GB;
      bool ready = ba.w <= 0 || scan_blocks(0, slab, ba.w);
      if (!ready && !didA) {
        // no A work left: every A item is being streamed by some warp; wait for the ones this item needs (bounded)
        unsigned spins = 0;
        while (!ready && !gave_up) {
          __nanosleep(100);
          ready = scan_blocks(0, slab, ba.w);
          if (++spins > (1u << 20) || ((spins & 255u) == 0 && ld_relaxed(fa.ctrl + kFlagOff) != 0)) {
            if (lane == 0) atomicExch(fa.ctrl + kFlagOff, 1);
            gave_up = true;
          }
        }
        ready = true;
      }
      if (ready) {
        asm volatile("fence.acquire.gpu;" ::: "memory");
        if (lane == 0) tb_raw = atomicAdd(fa.ctrl + kNextB, 1);   // the next B ticket: in flight while this one streams
        b_claimed = true;
        const int col0 = slab * fa.slabF;
        const int Fs = min(fa.slabF, fa.F - col0);
        if (fa.niso > 0) {   // this item's share of the vertices that no hyperedge touches
          const int i0 = (int)((int64_t)fa.niso * idx / GB), i1 = (int)((int64_t)fa.niso * (idx + 1) / GB);
          for (int i = i0 + sub; i < i1; i += G::kSub) {
            float *yp = fa.out[1] + (int64_t)__ldg(fa.iso + i) * fa.F + col0;
#pragma unroll
            for (int v = 0; v < VPL; ++v)
              if (sl * 4 + v * G::kStride < Fs) st_row_hint(yp + sl * 4 + v * G::kStride, make_float4(0.f, 0.f, 0.f, 0.f), pol_y);
          }
        }
        int32_t ps, pe;
        bounds(ba, bb, ps, pe);
        stream_run<1, SW, VPL, KV, HAS_WIN, PIPE>(fa, ps, pe, col0, Fs, sl, 0, pol_y);
        if (fa.track_b) publish(1, slab, idx);
        // discards: as in the split-role form -- the warp that finishes the last item of B block k makes ONE
        // non-blocking attempt at discard item k - doff
        if (GC > 0 && (idx % kBlk == kBlk - 1 || idx == GB - 1)) {
          const int ci = idx / kBlk - fa.doff;
          if (ci >= 0 && scan_blocks(1, slab, ci + 1)) {
            asm volatile("fence.acquire.gpu;" ::: "memory");
            const int4 ic = __ldg(tabC + 2 * ci);
            const int lines = Fs >> 5;
            const int totl = (ic.y - ic.x) * lines;
            for (int x = lane; x < totl; x += 32) {
              const int r = x / lines, l = x - r * lines;
              const int32_t e = __ldg(fa.dperm + ic.x + r);
              const float *p = fa.in[1] + (int64_t)e * fa.F + col0 + l * 32;
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
            }
          }
        }
        tb = -1;
        continue;
      }
    }
    if (!didA && tb < 0 && !b_claimed) {
      if (b_none) break;
    }
  }
}

template <int SW, int VPL, bool HAS_WIN>
int launch_alt(hgPlan *p, const FArgs &fa, bool pipe, int occ, cudaStream_t s) {
  // (occupancy, vectors in flight per lane): 3 x 8 or 2 x 16
  void (*kern)(const FArgs) = nullptr;
  if (occ <= 2) kern = astream_kernel<SW, VPL, 16, HAS_WIN, 2, true>;
  else kern = pipe ? astream_kernel<SW, VPL, 8, HAS_WIN, 3, true> : astream_kernel<SW, VPL, 8, HAS_WIN, 3, false>;
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
  if (per_sm < 1) per_sm = 1;
  if (occ < per_sm) per_sm = occ < 1 ? 1 : occ;
  const int64_t grid = (int64_t)p->sm_count * per_sm;
  kern<<<(unsigned)grid, kThreads, 0, s>>>(fa);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int dispatch_alt(hgPlan *p, const FArgs &fa, int sw, int vpl, bool pipe, bool has_win, int occ, cudaStream_t s) {
#define HG_CASE(SW_, VPL_)                                                                                   \
  if (sw == SW_ && vpl == VPL_)                                                                              \
    return has_win ? launch_alt<SW_, VPL_, true>(p, fa, pipe, occ, s) : launch_alt<SW_, VPL_, false>(p, fa, pipe, occ, s);
  HG_CASE(8, 1) HG_CASE(8, 2) HG_CASE(8, 4) HG_CASE(16, 1) HG_CASE(16, 2) HG_CASE(16, 4)
  HG_CASE(32, 1) HG_CASE(32, 2) HG_CASE(32, 4)
#undef HG_CASE
  return set_error(HG_EINVAL, "alternating stream: no kernel for sub-warp %d x %d vectors", sw, vpl);
}

template <int SW, int VPL, bool HAS_WIN>
int launch_split(hgPlan *p, const FArgs &fa, bool pipe, int occ, cudaStream_t s) {
  void (*kern)(const FArgs) = nullptr;
  if (occ <= 2) kern = pipe ? dstream_kernel<SW, VPL, HAS_WIN, 2, true> : dstream_kernel<SW, VPL, HAS_WIN, 2, false>;
  else kern = pipe ? dstream_kernel<SW, VPL, HAS_WIN, 3, true> : dstream_kernel<SW, VPL, HAS_WIN, 3, false>;
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
  if (per_sm < 1) per_sm = 1;
  if (occ < per_sm) per_sm = occ < 1 ? 1 : occ;
  int64_t grid = (int64_t)p->sm_count * per_sm;
  if (grid < 2) grid = 2;
  grid &= ~int64_t(1);                       // both roles get the same number of CTAs
  // every CTA must be resident: the two roles wait for each other
  kern<<<(unsigned)grid, kThreads, 0, s>>>(fa);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int dispatch_split(hgPlan *p, const FArgs &fa, int sw, int vpl, bool pipe, bool has_win, int occ, cudaStream_t s) {
#define HG_CASE(SW_, VPL_)                                                                                   \
  if (sw == SW_ && vpl == VPL_)                                                                              \
    return has_win ? launch_split<SW_, VPL_, true>(p, fa, pipe, occ, s) : launch_split<SW_, VPL_, false>(p, fa, pipe, occ, s);
  HG_CASE(8, 1) HG_CASE(8, 2) HG_CASE(8, 4) HG_CASE(16, 1) HG_CASE(16, 2) HG_CASE(16, 4)
  HG_CASE(32, 1) HG_CASE(32, 2) HG_CASE(32, 4)
#undef HG_CASE
  return set_error(HG_EINVAL, "split-role stream: no kernel for sub-warp %d x %d vectors", sw, vpl);
}

struct FCfg { int sw, vpl, occ, kv; bool pipe; };

template <int SW, int VPL, bool HAS_WIN>
int launch_one(hgPlan *p, const FArgs &fa, const FCfg &cfg, int ctas, cudaStream_t s) {
  // (occupancy, vectors in flight): 3 x 8 (pipelined or not), 2 x 16, 1 x 32
  void (*kern)(const FArgs) = nullptr;
  if (cfg.occ <= 1) kern = fstream_kernel<SW, VPL, 32, HAS_WIN, 1, true>;
  else if (cfg.occ == 2) kern = fstream_kernel<SW, VPL, 16, HAS_WIN, 2, true>;
  else kern = cfg.pipe ? fstream_kernel<SW, VPL, 8, HAS_WIN, 3, true> : fstream_kernel<SW, VPL, 8, HAS_WIN, 3, false>;
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
  if (per_sm < 1) per_sm = 1;
  if (cfg.occ < per_sm) per_sm = cfg.occ < 1 ? 1 : cfg.occ;
  if (ctas > 0 && ctas < per_sm) per_sm = ctas;
  int64_t grid = (int64_t)p->sm_count * per_sm;
  const int64_t useful = ceil_div<int64_t>((int64_t)fa.nitem * fa.nslab, kWarpsPerBlock);
  if (grid > useful) grid = useful;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, kThreads, 0, s>>>(fa);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

int dispatch(hgPlan *p, const FArgs &fa, const FCfg &cfg, bool has_win, int ctas, cudaStream_t s) {
#define HG_CASE(SW_, VPL_)                                                                             \
  if (cfg.sw == SW_ && cfg.vpl == VPL_)                                                                \
    return has_win ? launch_one<SW_, VPL_, true>(p, fa, cfg, ctas, s) : launch_one<SW_, VPL_, false>(p, fa, cfg, ctas, s);
  HG_CASE(8, 1) HG_CASE(8, 2) HG_CASE(8, 4) HG_CASE(16, 1) HG_CASE(16, 2) HG_CASE(16, 4)
  HG_CASE(32, 1) HG_CASE(32, 2) HG_CASE(32, 4)
#undef HG_CASE
  return set_error(HG_EINVAL, "fused stream: no kernel for sub-warp %d x %d vectors", cfg.sw, cfg.vpl);
}

__global__ void zero_rows_kernel(int64_t nrows, const int32_t *__restrict__ segs, const int32_t *__restrict__ seg_edge,
                                 float *__restrict__ xe, int F) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (w >= nrows) return;
  float *row = xe + (int64_t)seg_edge[segs[w]] * F;
  for (int c = lane; c < F; c += 32) row[c] = 0.0f;
}

}  // namespace

bool fstream_available(const hgPlan *plan, int F, bool force) {
  if (!plan->st_ready || F % 4 != 0) return false;
  if (force) return true;
  if (tune_get("fstream", 1) == 0) return false;
  // below ~64 MB of Y everything is L2-resident anyway and the two-pass form has the lower latency
  if ((double)plan->num_nodes * F * 4.0 < 64.0 * 1048576.0) return false;
  return plan->max_vdeg <= 65536;
}

int launch_fstream(hgPlan *p, const dev::Args &a, cudaStream_t s) {
  const int F = a.F;
  if (int rc = ensure_xe(p, F, s)) return rc;
  FCfg cfg{};
  // geometry: SW lanes x VPL 128-bit vectors per row slab; at least 8 lanes per row (<= 4 streams per warp)
  int slabF = F <= 512 ? F : 512;
  {
    const int slab_t = tune_get("fs_slab", 0);
    if (slab_t >= 32 && slab_t % 32 == 0 && slab_t < slabF) slabF = slab_t;
  }
  const int nslab = (F + slabF - 1) / slabF;
  cfg.sw = slabF <= 32 ? 8 : (slabF <= 64 ? 16 : 32);
  {
    const int sw_t = tune_get("fs_sw", 0);
    if (sw_t == 8 || sw_t == 16 || sw_t == 32) cfg.sw = sw_t;
    while (cfg.sw < 32 && slabF > cfg.sw * 16) cfg.sw *= 2;
  }
  cfg.vpl = slabF <= cfg.sw * 4 ? 1 : (slabF <= cfg.sw * 8 ? 2 : 4);
  const int ksub = 32 / cfg.sw;
  cfg.occ = tune_get("fs_occ", 3);
  if (cfg.occ > 3) cfg.occ = 3;
  if (cfg.occ < 1) cfg.occ = 1;
  cfg.pipe = tune_get("fs_pipe", cfg.vpl < 4 ? 1 : 0) != 0;
  const int ctas = tune_get("fs_ctas", 0);
  // item: ~16 KB of rows per warp by default
  int item_kb = tune_get("fs_item_kb", 16);
  int k0 = item_kb * 1024 / (slabF * 4 * ksub * kL0f);   // fine runs per sub-stream
  if (k0 < 1) k0 = 1;
  const int bpi = k0 * ksub;
  const int warps = p->sm_count * (ctas > 0 ? ctas : cfg.occ) * kWarpsPerBlock;
  int lagB = tune_get("fs_lag_b", -1), lagC = tune_get("fs_lag_c", -1);
  const int batch_t = 1;
  if (lagB < 0) lagB = warps * batch_t;   // every warp may hold a batch of unfinished items
  if (lagC < 0) lagC = warps * batch_t;
  // rows can be discarded line by line only if they are made of whole 128-byte lines
  const int discard = (F % 32 == 0 && tune_get("fs_discard", 1) != 0) ? 1 : 0;

  const int split_mode = tune_get("fs_split", 1);   // 0: merged ticket order, 1: split-role CTAs, 2: alternating warps
  const bool split = split_mode != 0;
  if (split) { lagB = 0; lagC = 0; }         // the merged order is not used: one cached table set per item size
  hgPlan::RingSched *sc = nullptr;
  if (int rc = fused_get_sched(p, bpi, lagB, lagC, nslab, discard, ksub, s, &sc)) return rc;
  if (p->nheavy_segs > 0) {
    zero_rows_kernel<<<(unsigned)ceil_div<int64_t>(p->nheavy_segs * 32, 256), 256, 0, s>>>(
        p->nheavy_segs, p->heavy_segs, p->seg_edge, p->xe, F);
    HG_CUDA_TRY(cudaGetLastError());
    ++p->kernels_launched;
  }
  HG_CUDA_TRY(cudaMemsetAsync(sc->ctrl, 0, ((size_t)kCntOff + (size_t)(sc->nblkA + sc->nblkB + 64) * nslab) * sizeof(int32_t), s));
  FArgs fa{};
  fa.src[0] = p->st_srcA; fa.dst[0] = p->st_dstA; fa.src[1] = p->st_srcB; fa.dst[1] = p->st_dstB;
  fa.in[0] = a.X; fa.out[0] = p->xe; fa.in[1] = p->xe; fa.out[1] = a.Y;
  fa.w_in = a.a_in;
  fa.w_o1[0] = a.s1; fa.w_o2[0] = a.s2; fa.w_o1[1] = a.a_out; fa.w_o2[1] = nullptr;
  fa.items = sc->items; fa.dperm = p->rg_dperm;
  fa.iso = p->st_perm + p->st_nunitB; fa.niso = (int32_t)p->st_niso;
  fa.ctrl = sc->ctrl;
  fa.nitem = sc->nitem; fa.nslab = nslab; fa.slabF = slabF; fa.F = F;
  fa.nblkA = sc->nblkA; fa.nblkB = sc->nblkB; fa.GA = sc->GA; fa.GB = sc->GB;
  fa.track_b = discard;
  // (X rows are gathered 1.76 times each on the bench graph and the L2 serves the repeats: no evict-first hint on X)
  fa.pol_x = tune_get("fs_pol_x", kPolNormal);
  fa.pol_xe_w = tune_get("fs_pol_xe_w", kPolLast);
  fa.pol_y = tune_get("fs_pol_y", kPolFirst);
  // (claiming and publishing tickets in batches of 2..32 was measured: no gain, and one configuration timed out on the
  //  Walmart shape -- profiles/r02_fstream_autonomous_warps_rejected.txt; the code path stays for the record, off)
  fa.batch = 1;
  fa.debug = tune_get("fs_debug", 0);
  p->rg_last_ctrl = sc->ctrl;
  ++p->kernels_launched;
  if (split) {
    fa.tabs = sc->tabs; fa.GC = sc->GC;
    const int extra = tune_get("fs_lead", -1);
    // stage A may run this many items ahead of the completed B prefix: what the graph needs + a margin of one
    // front of warps (so that neither role waits in the steady state)
    fa.maxlead = sc->lead + (extra >= 0 ? extra : warps / 2) + 2 * kBlk;
    const int doff_t = tune_get("fs_doff", -1);
    fa.doff = doff_t >= 0 ? doff_t : (warps / 2 + kBlk - 1) / kBlk + 8;   // ~ the B blocks in flight
    if (split_mode == 2) {
      if (doff_t < 0) fa.doff = (warps + kBlk - 1) / kBlk + 8;   // every warp may hold a B item
      return dispatch_alt(p, fa, cfg.sw, cfg.vpl, cfg.pipe, a.a_in != nullptr, cfg.occ, s);
    }
    return dispatch_split(p, fa, cfg.sw, cfg.vpl, cfg.pipe, a.a_in != nullptr, cfg.occ, s);
  }
  return dispatch(p, fa, cfg, a.a_in != nullptr, ctas, s);
}

}  // namespace hg
