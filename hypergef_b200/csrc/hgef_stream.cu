// hgef_stream.cu -- the STREAM form of the fused aggregation: both stages as lean, register-only row streams
// (the default for graphs whose Y exceeds the L2).
//
// What the profiles of the earlier forms said (profiles/r01_ncu_prof_r1_pull_f128.txt): the gather-only
// two-phase form has ideal DRAM traffic but spends ~45 warp instructions per gathered row (tile staging in
// shared memory, cp.async ring, per-unit bookkeeping for units of 2-4 rows) at 16 resident warps per SM.
//
// Here everything a row needs is precomputed once per graph into a ROW PROGRAM (hg_plan_create):
//     src[p]  row to gather at position p
//     dst[p]  output row of the unit that owns p; bit31 = "last member of its unit" (store now),
//             bit30 = "heavy: reduce into the pre-zeroed row, do not store"
//     run[r]  first position of base run r (unit-aligned, ~kL0 positions each)
// for stage A (units = balancer segments of H^T, gather X, output Xe) and stage B (units = vertices ordered
// by their last hyperedge, gather Xe, output Y).  Warps claim items of a few runs from one counter.  A sub-warp
// of SW lanes streams one run: the src / dst words of SW positions are held one per lane (the next chunk's
// are prefetched) and handed out by shuffles; rows are loaded in half-batches into two alternating register
// sets, half h + 1 issued before half h is consumed, so every lane has 4..8 row vectors in flight; a unit end
// is one uniform test of a ballot mask and costs two shuffles, a scale and one 128-bit store per lane.  No
// shared memory, no barriers, no atomics but the ticket (and red.v4 for heavy hyperedges).  Per-stage kernels
// take every array base from the constant bank: 64 registers at one vector per lane.
// What bounds it: DRAM at >= 1 KB rows (94-96 % of the copy peak, 1.3x the algorithmic
// traffic because Xe makes a round trip); below that, how much the L2 serves at 24-32 warps per SM.
//
// One launch per stage.  (Merging both stages into one persistent launch so that Xe is handed over inside the L2
// was built six ways this round -- hgef_fstream.cu, hgef_ring.cu, lab library -- and is slower: DESIGN.md
// section 4.)  Wide rows (> 512 floats) are processed as column SLABS.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cstdlib>

#include "hgef_stream.cuh"

namespace hg {
namespace {
using namespace dev;

constexpr int kVec = 8;      // 128-bit row loads in flight per lane

struct StreamArgs {
  const int32_t *src[2], *dst[2], *run[2];   // row programs: [0] stage A, [1] stage B
  const float *in[2];
  float *out[2];
  const float *w_in[2], *w_o1[2], *w_o2[2];  // gather-side weight per source row; output scales per output row
  int32_t nrun0[2];                          // base runs
  int32_t *ctrl;
  const int32_t *iso;                        // vertices in no hyperedge: Y row = 0
  int32_t niso;
  int32_t nitem;                             // tickets per slab
  int32_t nslab, slabF;                      // column slabs of slabF floats
  int32_t F;                                 // row stride (floats)
  int32_t k0;                                // base runs per sub-warp run
  int32_t stage;                             // the stage this launch runs
  int32_t y_stream;                          // stage-B output stores carry the streaming (evict-first) hint
  int32_t pdl;                               // stage B is a programmatic dependent launch of stage A
};

__device__ __forceinline__ float4 ld_row16(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st_row16(float *p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_row16_stream(float *p, float4 v) {   // written once, never re-read
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int SW, int VPL>
struct SGeo {
  static constexpr int kSub = 32 / SW;                          // row streams per warp
  static constexpr int kU = (kVec / VPL) < SW ? (kVec / VPL) : SW;   // rows in flight per stream
  static constexpr int kHB = kU >= 2 ? kU / 2 : 1;              // rows per half-batch (two register sets alternate)
  static constexpr int kNH = SW / kHB;                          // half-batches per chunk of SW positions (even)
  static constexpr int kStride = SW * 4;                        // floats between a lane's vectors
};

// STAGE 0 / 1: the stage this launch runs (its arrays are addressed straight from the constant bank).
// PIPE: rows are loaded in half-batches into two alternating register sets (half h + 1 is issued before half
// h is consumed); otherwise whole batches are loaded and then consumed (better for the widest rows, where
// a half-batch is a single 2 KB row).
template <int SW, int VPL, bool HAS_WIN, int STAGE, int MINB, bool PIPE>
__global__ void __launch_bounds__(kThreads, MINB) stream_kernel(const StreamArgs sa) {
  using G = SGeo<SW, VPL>;
  constexpr int HB = PIPE ? G::kHB : G::kU, NH = SW / HB;
  static_assert(!PIPE || (NH >= 2 && NH % 2 == 0), "a chunk is a whole number of half-batch pairs");
  const int lane = threadIdx.x & 31;
  const int sub = lane / SW, sl = lane % SW;
  const int col = sl * 4;
  const int F = sa.F;
  const uint32_t row_bytes = (uint32_t)F * 4u;
  const int total = sa.nitem * sa.nslab;
  uint32_t pat = 0;   // bit (stream * SW) for every row stream of the warp
#pragma unroll
  for (int q = 0; q < G::kSub; ++q) pat |= 1u << (q * SW);

  if (sa.pdl) {
    // stage A lets its dependent (stage B) be scheduled as soon as every A CTA has started; stage B must not
    // touch Xe before stage A's memory operations are complete and flushed
    if (STAGE == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    else asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(sa.ctrl, 1);
    t = __shfl_sync(kFull, t, 0);
    if (t >= total) break;
    const int slab = sa.nslab > 1 ? t / sa.nitem : 0;
    const int k = t - slab * sa.nitem;
    const int col0 = slab * sa.slabF;
    const int Fs = min(sa.slabF, F - col0);
    constexpr int stage = STAGE;
    const int gid = k;
    const float *in = sa.in[stage] + col0;
    float *out = sa.out[stage] + col0;
    // column mask of this lane's vectors; a masked vector LOADS column 0 of the slab instead (no branch
    // around the load) and is never stored
    bool ok[VPL];
    int off[VPL];
    const char *in_v[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
      ok[v] = col + v * G::kStride < Fs;
      off[v] = ok[v] ? col + v * G::kStride : 0;
      in_v[v] = reinterpret_cast<const char *>(in + off[v]);
    }

    // this ticket's share of the vertices that no hyperedge touches (only the B side has any)
    if (sa.niso > 0 && STAGE == 1) {
      const int i0 = (int)((int64_t)sa.niso * k / sa.nitem), i1 = (int)((int64_t)sa.niso * (k + 1) / sa.nitem);
      for (int i = i0 + sub; i < i1; i += G::kSub) {
        float *yp = sa.out[1] + (int64_t)__ldg(sa.iso + i) * F + col0;
#pragma unroll
        for (int v = 0; v < VPL; ++v)
          if (ok[v]) st_row16_stream(yp + off[v], make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }

    const int32_t *__restrict__ src = sa.src[stage];
    const int32_t *__restrict__ dst = sa.dst[stage];
    const float *__restrict__ w_in = sa.w_in[stage];
    const float *__restrict__ w_o1 = sa.w_o1[stage];
    const float *__restrict__ w_o2 = sa.w_o2[stage];
    const int nrun0 = sa.nrun0[stage];
    const int64_t r0 = ((int64_t)gid * G::kSub + sub) * sa.k0;
    const int32_t ps = __ldg(sa.run[stage] + min(r0, (int64_t)nrun0));
    const int32_t pe = __ldg(sa.run[stage] + min(r0 + sa.k0, (int64_t)nrun0));

    // ---- stream the run.  Chunks of SW positions: their src / dst words one per lane, handed out by
    // shuffles.  Rows are loaded in half-batches of HB rows into two alternating register sets: half
    // h + 1 is issued before half h is consumed, so a stream always has HB..2HB row loads in flight.
    float4 acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int32_t cb = ps;
    uint32_t c_src = 0, c_dst = 0;
    if (cb + sl < pe) {
      c_src = (uint32_t)__ldg(src + cb + sl);
      c_dst = (uint32_t)__ldg(dst + cb + sl);
    }
    float4 x0[HB][VPL], x1[HB][VPL];
    // (positions past the end of the run carry src word 0: row 0 is loaded and never used for output --
    //  a run ends with a unit end, which resets `acc`, and `acc` restarts from zero with every item)
    auto load_half = [&](float4(&x)[HB][VPL], uint32_t words, int j0) {
#pragma unroll
      for (int u = 0; u < HB; ++u) {
        const uint32_t id = __shfl_sync(kFull, words, j0 + u, SW);
        const uint64_t rb = (uint64_t)id * row_bytes;   // one IMAD.WIDE.U32 per vector below
#pragma unroll
        for (int v = 0; v < VPL; ++v) x[u][v] = ld_row16(reinterpret_cast<const float *>(in_v[v] + rb));
      }
    };
    if (PIPE) load_half(x0, c_src, 0);
    while (__any_sync(kFull, cb < pe)) {
      float c_w = 1.0f, c_sc = 1.0f;
      if (cb + sl < pe) {
        if (HAS_WIN && w_in) c_w = __ldg(w_in + c_src);
        if (c_dst & kEnd) {
          const uint32_t orow = c_dst & kRowMask;
          if (w_o1) c_sc = __ldg(w_o1 + orow);
          if (w_o2) c_sc *= __ldg(w_o2 + orow);
        }
      }
      // unit-end flags of the chunk: bit (stream * SW + position); a step whose position ends no unit in
      // any stream of the warp skips the output bookkeeping with one uniform test
      const uint32_t endm = __ballot_sync(kFull, (c_dst & kEnd) != 0);
      uint32_t n_src = 0, n_dst = 0;   // next chunk's words: in flight while this chunk streams
      if (cb + SW + sl < pe) {
        n_src = (uint32_t)__ldg(src + cb + SW + sl);
        n_dst = (uint32_t)__ldg(dst + cb + SW + sl);
      }
      auto consume_half = [&](const float4(&x)[HB][VPL], int j0) {
#pragma unroll
        for (int u = 0; u < HB; ++u) {
          const int j = j0 + u;
          float w = 1.0f;
          if (HAS_WIN) w = __shfl_sync(kFull, c_w, j, SW);
#pragma unroll
          for (int v = 0; v < VPL; ++v) {
            if (HAS_WIN) {
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
                  if (d & kHeavy) red_add_v4(o, r);
                  else if (stage == 1 && sa.y_stream) st_row16_stream(o, r);
                  else st_row16(o, r);
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
  // the last warp of the grid to leave resets the ticket counter for the next launch (ctrl[2] counts leavers),
  // so a call needs no memset between its launches
  if (lane == 0) {
    const int nw = (int)(gridDim.x * (blockDim.x >> 5));
    if (atomicAdd(sa.ctrl + 2, 1) == nw - 1) {
      sa.ctrl[0] = 0;
      sa.ctrl[2] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------
// Row-program construction (once per graph)
// ------------------------------------------------------------------------------------------
#define GRID(n) (unsigned)ceil_div<int64_t>((n), 256), 256

// one warp per unit: src = gathered row (| END on the unit's last position), dst = output row (| HEAVY)
__global__ void prog_fill_kernel(int64_t nunit, const int32_t *__restrict__ ptr_in,    // where the unit's members are
                                 const int32_t *__restrict__ ptr_out,                  // where they go (may alias)
                                 const int32_t *__restrict__ unit_of,                  // optional: source unit of slot i
                                 const int32_t *__restrict__ ind, const int32_t *__restrict__ out_row,
                                 const int32_t *__restrict__ heavy_slot, const int32_t *__restrict__ unit_need,
                                 int32_t *__restrict__ src, int32_t *__restrict__ dst, int32_t *__restrict__ need) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= nunit) return;
  const int32_t u = unit_of ? unit_of[i] : (int32_t)i;
  const int32_t a = ptr_in[u], n = ptr_in[u + 1] - a, o = ptr_out[i];
  uint32_t d = (uint32_t)(out_row ? out_row[u] : u);
  if (heavy_slot && heavy_slot[u] >= 0) d |= kHeavy;
  const int32_t nd = unit_need ? unit_need[i] : 0;
  for (int32_t j = lane; j < n; j += 32) {
    src[o + j] = ind[a + j];
    dst[o + j] = (int32_t)(j == n - 1 ? d | kEnd : d);
    if (need) need[o + j] = nd;
  }
}

// run[r] = smallest unit start >= r * L0 (units are never cut); run[nrun] = npos
__global__ void run_ptr_kernel(int64_t nrun, int64_t nunit, const int32_t *__restrict__ ptr, int64_t npos,
                               int32_t *__restrict__ run, int L0) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > nrun) return;
  if (r == nrun) { run[r] = (int32_t)npos; return; }
  const int64_t target = r * L0;
  int64_t lo = 0, hi = nunit;          // first unit index with ptr[idx] >= target (ptr[nunit] = npos >= target)
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (ptr[mid] < target) lo = mid + 1; else hi = mid;
  }
  run[r] = ptr[lo];
}

// last position (exclusive) of every hyperedge in the stage-A order
__global__ void edge_end_kernel(int64_t nseg, const int32_t *__restrict__ key, const int32_t *__restrict__ seg_edge,
                                int32_t *__restrict__ edge_end) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  if (s + 1 == nseg || seg_edge[s + 1] != seg_edge[s]) edge_end[seg_edge[s]] = key[s + 1];
}

// sort key of a vertex: its last hyperedge (isolated vertices: M, they go to the end)
__global__ void vertex_key_kernel(int64_t n, const int32_t *__restrict__ h_ptr, const int32_t *__restrict__ h_ind,
                                  int32_t m, int32_t *__restrict__ keys, int32_t *__restrict__ ids) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  keys[v] = h_ptr[v + 1] > h_ptr[v] ? h_ind[h_ptr[v + 1] - 1] : m;
  ids[v] = (int32_t)v;
}

__global__ void perm_deg_kernel(int64_t nb, const int32_t *__restrict__ perm, const int32_t *__restrict__ keys_sorted,
                                const int32_t *__restrict__ h_ptr, const int32_t *__restrict__ edge_end,
                                int32_t *__restrict__ deg, int32_t *__restrict__ unit_need) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > nb) return;
  if (i == nb) { deg[i] = 0; return; }
  const int32_t v = perm[i];
  deg[i] = h_ptr[v + 1] - h_ptr[v];
  unit_need[i] = edge_end[keys_sorted[i]];
}

// number of vertices with at least one hyperedge = index after the last sorted key < M
__global__ void count_units_kernel(int64_t n, const int32_t *__restrict__ keys_sorted, int32_t m, int32_t *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n && keys_sorted[i] < m && (i + 1 == n || keys_sorted[i + 1] >= m)) *out = (int32_t)(i + 1);
}

__global__ void zero_rows_kernel(int64_t nrows, const int32_t *__restrict__ segs, const int32_t *__restrict__ seg_edge,
                                 float *__restrict__ xe, int F) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (w >= nrows) return;
  float *row = xe + (int64_t)seg_edge[segs[w]] * F;
  for (int c = lane; c < F; c += 32) row[c] = 0.0f;
}

template <typename T>
int dev_alloc(T **p, size_t n) {
  if (cudaMalloc((void **)p, (n ? n : 1) * sizeof(T)) != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ENOMEM, "stream plan: cannot allocate %zu bytes", n * sizeof(T));
  }
  return HG_OK;
}

}  // namespace

void stream_free(hgPlan *p) {
  cudaFree(p->st_srcA); cudaFree(p->st_dstA); cudaFree(p->st_runA);
  cudaFree(p->st_srcB); cudaFree(p->st_dstB); cudaFree(p->st_needB); cudaFree(p->st_runB);
  cudaFree(p->st_perm); cudaFree(p->st_ctrl); cudaFree(p->st_ptrB);
}

// Builds the two row programs.  Needs the canonical segment schedule and H (build_pull).
int build_stream(hgPlan *p, cudaStream_t s) {
  const int64_t N = p->num_nodes, M = p->num_edges, Z = p->nnz, S = p->nseg;
  if (!p->canonical || p->h_ptr == nullptr || Z == 0) return HG_OK;
  if (N >= (int64_t(1) << 30) || M >= (int64_t(1) << 30)) return HG_OK;   // row ids share a word with two flags
  {   // the segments must tile [0, Z) exactly (the balancer's do)
    int32_t k0 = -1, kS = -1;
    HG_CUDA_TRY(cudaMemcpyAsync(&k0, p->key, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    HG_CUDA_TRY(cudaMemcpyAsync(&kS, p->key + S, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    HG_CUDA_TRY(cudaStreamSynchronize(s));
    if (k0 != 0 || kS != Z) return HG_OK;
  }
  // stage A: units = balancer segments in order
  if (int rc = dev_alloc(&p->st_srcA, Z)) return rc;
  if (int rc = dev_alloc(&p->st_dstA, Z)) return rc;
  p->st_nrunA = ceil_div<int64_t>(Z, kL0);
  if (int rc = dev_alloc(&p->st_runA, p->st_nrunA + 1)) return rc;
  prog_fill_kernel<<<(unsigned)ceil_div<int64_t>(S * 32, 256), 256, 0, s>>>(
      S, p->key, p->key, nullptr, p->colind, p->seg_edge, p->seg_slot, nullptr, p->st_srcA, p->st_dstA, nullptr);
  run_ptr_kernel<<<GRID(p->st_nrunA + 1), 0, s>>>(p->st_nrunA, S, p->key, Z, p->st_runA, kL0);
  HG_CUDA_TRY(cudaGetLastError());

  // stage B: units = vertices, ordered by their last hyperedge (the stage-A position they wait for)
  DevBuf<int32_t> keys, ids, keys_s, edge_end, deg, unit_need;
  HG_CUDA_TRY(keys.alloc(N)); HG_CUDA_TRY(ids.alloc(N)); HG_CUDA_TRY(keys_s.alloc(N));
  HG_CUDA_TRY(edge_end.alloc(M + 1)); HG_CUDA_TRY(deg.alloc(N + 1));
  if (int rc = dev_alloc(&p->st_ptrB, N + 1)) return rc;
  HG_CUDA_TRY(unit_need.alloc(N));
  if (int rc = dev_alloc(&p->st_perm, N)) return rc;
  HG_CUDA_TRY(cudaMemsetAsync(edge_end.p, 0, (size_t)(M + 1) * sizeof(int32_t), s));
  edge_end_kernel<<<GRID(S), 0, s>>>(S, p->key, p->seg_edge, edge_end.p);
  vertex_key_kernel<<<GRID(N), 0, s>>>(N, p->h_ptr, p->h_ind, (int32_t)M, keys.p, ids.p);
  HG_CUDA_TRY(cudaGetLastError());
  int end_bit = 1;
  while (end_bit < 32 && (int64_t(1) << end_bit) <= M) ++end_bit;
  size_t bytes = 0;
  HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys.p, keys_s.p, ids.p, p->st_perm, N, 0, end_bit, s));
  {
    DevBuf<char> ws;
    HG_CUDA_TRY(ws.alloc(bytes));
    HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws.p, bytes, keys.p, keys_s.p, ids.p, p->st_perm, N, 0, end_bit, s));
    HG_CUDA_TRY(cudaStreamSynchronize(s));
  }
  perm_deg_kernel<<<GRID(N + 1), 0, s>>>(N, p->st_perm, keys_s.p, p->h_ptr, edge_end.p, deg.p, unit_need.p);
  HG_CUDA_TRY(cudaGetLastError());
  {
    size_t b2 = 0;
    HG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, b2, deg.p, p->st_ptrB, N + 1, s));
    DevBuf<char> ws;
    HG_CUDA_TRY(ws.alloc(b2));
    HG_CUDA_TRY(cub::DeviceScan::ExclusiveSum(ws.p, b2, deg.p, p->st_ptrB, N + 1, s));
    HG_CUDA_TRY(cudaStreamSynchronize(s));
  }
  // isolated vertices are the tail of the permutation (sorted key == M)
  {
    DevBuf<int32_t> nb;
    HG_CUDA_TRY(nb.alloc(1));
    HG_CUDA_TRY(cudaMemsetAsync(nb.p, 0, sizeof(int32_t), s));
    count_units_kernel<<<GRID(N), 0, s>>>(N, keys_s.p, (int32_t)M, nb.p);
    int32_t h_nb = 0;
    HG_CUDA_TRY(cudaMemcpyAsync(&h_nb, nb.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    HG_CUDA_TRY(cudaStreamSynchronize(s));
    p->st_nunitB = h_nb;
    p->st_niso = N - h_nb;
  }
  if (int rc = dev_alloc(&p->st_srcB, Z)) return rc;
  if (int rc = dev_alloc(&p->st_dstB, Z)) return rc;
  if (int rc = dev_alloc(&p->st_needB, Z)) return rc;
  p->st_nrunB = ceil_div<int64_t>(Z, kL0);
  if (int rc = dev_alloc(&p->st_runB, p->st_nrunB + 1)) return rc;
  if (int rc = dev_alloc(&p->st_ctrl, 2 * kCtrlHdr)) return rc;
  HG_CUDA_TRY(cudaMemsetAsync(p->st_ctrl, 0, 2 * kCtrlHdr * sizeof(int32_t), s));   // the launches reset it themselves
  const int64_t NB = p->st_nunitB;
  if (NB > 0)
    prog_fill_kernel<<<(unsigned)ceil_div<int64_t>(NB * 32, 256), 256, 0, s>>>(
        NB, p->h_ptr, p->st_ptrB, p->st_perm, p->h_ind, nullptr, nullptr, unit_need.p, p->st_srcB, p->st_dstB, p->st_needB);
  run_ptr_kernel<<<GRID(p->st_nrunB + 1), 0, s>>>(p->st_nrunB, NB, p->st_ptrB, Z, p->st_runB, kL0);
  HG_CUDA_TRY(cudaGetLastError());
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  p->st_ready = 1;
  return HG_OK;
}

// unit-aligned run tables of both stages for runs of L0 positions (the fused forms use finer runs than kL0)
int stream_build_runs(hgPlan *p, int L0, int32_t **runA, int64_t *nrunA, int32_t **runB, int64_t *nrunB, cudaStream_t s) {
  const int64_t Z = p->nnz;
  *nrunA = *nrunB = ceil_div<int64_t>(Z, L0);
  if (int rc = dev_alloc(runA, *nrunA + 1)) return rc;
  if (int rc = dev_alloc(runB, *nrunB + 1)) return rc;
  run_ptr_kernel<<<GRID(*nrunA + 1), 0, s>>>(*nrunA, p->nseg, p->key, Z, *runA, L0);
  run_ptr_kernel<<<GRID(*nrunB + 1), 0, s>>>(*nrunB, p->st_nunitB, p->st_ptrB, Z, *runB, L0);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

namespace {

struct StreamCfg {
  int sw, vpl, slabF, nslab, k0, ctas, occ;
  bool pipe, pdl;
};

template <int SW, int VPL, bool HAS_WIN, int STAGE>
int launch_one(hgPlan *p, StreamArgs &sa, const StreamCfg &cfg, cudaStream_t s) {
  auto kern = cfg.pipe ? (cfg.occ == 4 ? stream_kernel<SW, VPL, HAS_WIN, STAGE, 4, true> : stream_kernel<SW, VPL, HAS_WIN, STAGE, 3, true>)
                       : stream_kernel<SW, VPL, HAS_WIN, STAGE, 3, false>;
  int per_sm = 0;
  HG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
  if (per_sm < 1) per_sm = 1;
  if (cfg.ctas > 0 && cfg.ctas < per_sm) per_sm = cfg.ctas;
  int64_t grid = (int64_t)p->sm_count * per_sm;
  const int64_t useful = ceil_div<int64_t>((int64_t)sa.nitem * sa.nslab, kWarpsPerBlock);
  if (grid > useful) grid = useful;
  if (grid < 1) grid = 1;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(kThreads);
  lc.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = (STAGE == 1 && sa.pdl) ? 1 : 0;
  HG_CUDA_TRY(cudaLaunchKernelEx(&lc, kern, sa));
  return HG_OK;
}

int dispatch(hgPlan *p, StreamArgs &sa, const StreamCfg &cfg, int stage, bool has_win, cudaStream_t s) {
#define HG_CASE(SW_, VPL_)                                                                             \
  if (cfg.sw == SW_ && cfg.vpl == VPL_) {                                                              \
    if (stage == 1) return launch_one<SW_, VPL_, false, 1>(p, sa, cfg, s);                             \
    return has_win ? launch_one<SW_, VPL_, true, 0>(p, sa, cfg, s) : launch_one<SW_, VPL_, false, 0>(p, sa, cfg, s);    \
  }
  HG_CASE(4, 1) HG_CASE(4, 2) HG_CASE(4, 4) HG_CASE(8, 1) HG_CASE(8, 2) HG_CASE(8, 4)
  HG_CASE(16, 1) HG_CASE(16, 2) HG_CASE(16, 4) HG_CASE(32, 1) HG_CASE(32, 2) HG_CASE(32, 4)
#undef HG_CASE
  return set_error(HG_EINVAL, "stream: no kernel for sub-warp %d x %d vectors", cfg.sw, cfg.vpl);
}

}  // namespace

bool stream_available(const hgPlan *plan, int F, bool force) {
  if (!plan->st_ready) return false;
  if (force) return true;
  if (tune_get("stream", 1) == 0) return false;
  // below ~64 MB of Y the two-pass form (everything L2-resident, one warp per segment) has the lower latency
  if ((double)plan->num_nodes * F * 4.0 < (double)tune_get("stream_min_mb", 64) * 1048576.0) return false;
  return plan->max_vdeg <= 65536;   // one sub-warp walks a vertex's hyperedges
}

int ensure_xe(hgPlan *plan, int F, cudaStream_t s) {
  return plan_grow(plan, &plan->xe, &plan->xe_floats, (size_t)plan->num_edges * F, s, "hyperedge features");
}

// Launch geometry.  Defaults from the sweeps in profiles/; every one can be overridden through hg_tune_set
// (st_slab, st_sw, st_l, st_ctas, st_occ, st_pipe, st_cs, st_pdl, st_only) -- read from a table, not the environment.
int launch_stream_stages(hgPlan *p, const dev::Args &a, int stages, cudaStream_t s, float *xe_ext) {
  const int F = a.F;
  // the hyperedge features: the plan's own buffer, or the caller's (hg_plan_edge_reduce / hg_plan_edge_scatter)
  if (!xe_ext)
    if (int rc = ensure_xe(p, F, s)) return rc;
  float *const xe = xe_ext ? xe_ext : p->xe;
  StreamCfg cfg{};
  // geometry: SW lanes x VPL 128-bit vectors per row slab
  int slabF = tune_get("st_slab", 0);
  if (slabF <= 0) slabF = F <= 512 ? F : 512;
  if (slabF > 512) slabF = 512;
  slabF = (slabF + 3) / 4 * 4;
  if (slabF > F) slabF = F;
  cfg.slabF = slabF;
  cfg.nslab = (F + slabF - 1) / slabF;
  cfg.sw = slabF <= 16 ? 4 : (slabF <= 32 ? 8 : (slabF <= 64 ? 16 : 32));
  {   // st_sw: lanes per row (the lane then holds 1, 2 or 4 vectors of the row)
    const int sw_t = tune_get("st_sw", 0);
    if (sw_t == 4 || sw_t == 8 || sw_t == 16 || sw_t == 32) cfg.sw = sw_t;
    while (cfg.sw < 32 && slabF > cfg.sw * 16) cfg.sw *= 2;
  }
  cfg.vpl = slabF <= cfg.sw * 4 ? 1 : (slabF <= cfg.sw * 8 ? 2 : 4);
  const int ksub = 32 / cfg.sw;
  // run length per row stream: ~32 KB of gathered rows per warp item by default
  int L = tune_get("st_l", 0);
  if (L <= 0) {
    L = (32 * 1024) / (slabF * 4 * ksub);
    if (L < kL0) L = kL0;
    if (L > 256) L = 256;
  }
  cfg.k0 = (L + kL0 - 1) / kL0;
  cfg.ctas = tune_get("st_ctas", 0);
  cfg.occ = tune_get("st_occ", 3);
  cfg.pipe = tune_get("st_pipe", cfg.vpl < 4 ? 1 : 0) != 0;
  cfg.pdl = tune_get("st_pdl", 1) != 0;
  const int bpi = cfg.k0 * ksub;
  const bool has_win = a.a_in != nullptr;

  if (p->nheavy_segs > 0 && (stages & 1)) {
    zero_rows_kernel<<<(unsigned)ceil_div<int64_t>(p->nheavy_segs * 32, 256), 256, 0, s>>>(
        p->nheavy_segs, p->heavy_segs, p->seg_edge, xe, F);
    HG_CUDA_TRY(cudaGetLastError());
    ++p->kernels_launched;
  }
  StreamArgs sa{};
  sa.src[0] = p->st_srcA; sa.dst[0] = p->st_dstA; sa.run[0] = p->st_runA; sa.nrun0[0] = (int32_t)p->st_nrunA;
  sa.src[1] = p->st_srcB; sa.dst[1] = p->st_dstB; sa.run[1] = p->st_runB; sa.nrun0[1] = (int32_t)p->st_nrunB;
  sa.in[0] = a.X; sa.out[0] = xe; sa.w_in[0] = a.a_in; sa.w_o1[0] = a.s1; sa.w_o2[0] = a.s2;
  sa.in[1] = xe; sa.out[1] = a.Y; sa.w_in[1] = nullptr; sa.w_o1[1] = a.a_out; sa.w_o2[1] = nullptr;
  sa.iso = p->st_perm + p->st_nunitB; sa.niso = (int32_t)p->st_niso;
  sa.nslab = cfg.nslab; sa.slabF = cfg.slabF; sa.F = F; sa.k0 = cfg.k0;
  sa.y_stream = tune_get("st_cs", 1);
  const int32_t GA = (int32_t)ceil_div<int64_t>(p->st_nrunA, bpi), GB = (int32_t)ceil_div<int64_t>(p->st_nrunB, bpi);

  // two launches, each stage its own ticket counter; the counters reset themselves (the last warp to leave a
  // launch zeroes its counter), so a call is exactly two kernel launches and stage B can be a programmatic
  // dependent launch: its CTAs are scheduled while stage A drains and wait (griddepcontrol.wait) for A's
  // memory to be flushed before they touch Xe
  const int only = tune_get("st_only", 0);   // timing: 1 = stage A, 2 = stage B
  if (only == 1 || only == 2) stages &= only;
  if (stages & 1) {
    sa.stage = 0; sa.nitem = GA; sa.ctrl = p->st_ctrl; sa.pdl = (cfg.pdl && stages == 3) ? 1 : 0;
    ++p->kernels_launched;
    if (int rc = dispatch(p, sa, cfg, 0, has_win, s)) return rc;
  }
  if (!(stages & 2)) return HG_OK;
  sa.stage = 1; sa.nitem = GB; sa.ctrl = p->st_ctrl + kCtrlHdr; sa.pdl = (cfg.pdl && stages == 3) ? 1 : 0;
  ++p->kernels_launched;
  return dispatch(p, sa, cfg, 1, false, s);
}

int launch_stream(hgPlan *p, const dev::Args &a, cudaStream_t s) { return launch_stream_stages(p, a, 3, s, nullptr); }

// ---- feature lengths that are not a multiple of 4: the same kernels on rows padded to the next multiple
namespace {
__global__ void pad_rows_kernel(int64_t n, int F, int Fp, const float *__restrict__ in, float *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // one thread per padded element
  if (i >= n * Fp) return;
  const int64_t r = i / Fp;
  const int c = (int)(i - r * Fp);
  out[i] = c < F ? in[r * F + c] : 0.0f;
}
__global__ void unpad_rows_kernel(int64_t n, int F, int Fp, const float *__restrict__ in, float *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;   // one thread per output element
  if (i >= n * F) return;
  const int64_t r = i / F;
  const int c = (int)(i - r * F);
  out[i] = in[r * Fp + c];
}
}  // namespace

int launch_stream_padded(hgPlan *p, const dev::Args &a, cudaStream_t s) {
  const int F = a.F, Fp = (F + 3) / 4 * 4;
  const int64_t N = p->num_nodes;
  if (int rc = plan_grow(p, &p->pad_x, &p->pad_x_floats, (size_t)N * Fp, s, "padded input")) return rc;
  if (int rc = plan_grow(p, &p->pad_y, &p->pad_y_floats, (size_t)N * Fp, s, "padded output")) return rc;
  pad_rows_kernel<<<GRID(N * Fp), 0, s>>>(N, F, Fp, a.X, p->pad_x);
  HG_CUDA_TRY(cudaGetLastError());
  dev::Args b = a;
  b.X = p->pad_x; b.Y = p->pad_y; b.F = Fp;
  if (int rc = launch_stream(p, b, s)) return rc;
  unpad_rows_kernel<<<GRID(N * F), 0, s>>>(N, F, Fp, p->pad_y, a.Y);
  HG_CUDA_TRY(cudaGetLastError());
  p->kernels_launched += 2;
  return HG_OK;
}

}  // namespace hg
