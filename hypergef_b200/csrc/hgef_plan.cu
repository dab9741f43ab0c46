// hgef_plan.cu -- derive, once, what the fused kernel needs from the balancer output.
//
// The reference kernel (hgnnaggr_cuda.cu:14-47) is launched over GROUPS: a hyperedge cut
// into w segments owns w*w groups, so each of its member rows is gathered w times.  The
// plan turns the same arrays into a schedule over SEGMENTS (every member row read once):
//   seg_edge[s]  hyperedge of segment s            (from the diagonal groups st == ed)
//   seg_slot[s]  -1 when the hyperedge has one segment, else the scratch row that its
//                segments reduce into before the scatter
// and verifies that the group arrays are exactly the balancer's full cross product
// (`canonical`); if they are not, hg_aggr_forward falls back to the literal group
// schedule, so arbitrary user-built group arrays keep the reference's meaning.
// Every index is range-checked here (HG_EGRAPH), so the hot kernels carry no checks.
#include <cub/device/device_scan.cuh>

#include <climits>

#include "hgef_plan.cuh"

namespace hg {
int build_stream(hgPlan *p, cudaStream_t s);
void stream_free(hgPlan *p);
void ring_free(hgPlan *p);
namespace {

enum : int { kBadIndex = 1, kBadKey = 2, kBadGroup = 4, kNotCanonical = 8 };

__global__ void check_colind_kernel(int64_t nnz, const int32_t *__restrict__ colind, int64_t n,
                                    int *__restrict__ flags) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < nnz && (colind[p] < 0 || colind[p] >= n)) atomicOr(flags, kBadIndex);
}

__global__ void check_key_kernel(int64_t nseg, const int32_t *__restrict__ key, int64_t nnz,
                                 int *__restrict__ flags, int *__restrict__ max_len) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  int32_t a = key[s], b = key[s + 1];
  if (a < 0 || b < a || b > nnz) atomicOr(flags, kBadKey);
  else atomicMax(max_len, b - a);
}

__global__ void diag_groups_kernel(int64_t ngroup, const int32_t *__restrict__ row,
                                   const int32_t *__restrict__ st, const int32_t *__restrict__ ed,
                                   int64_t nseg, int64_t nedge, int32_t *__restrict__ seg_edge,
                                   int *__restrict__ flags) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= ngroup) return;
  int32_t a = st[g], b = ed[g], r = row[g];
  if (a < 0 || a >= nseg || b < 0 || b >= nseg || r < 0 || r >= nedge) {
    atomicOr(flags, kBadGroup);
    return;
  }
  if (a == b) seg_edge[a] = r;
}

// first[s] = 1 when segment s opens a new hyperedge
__global__ void seg_first_kernel(int64_t nseg, const int32_t *__restrict__ seg_edge,
                                 int32_t *__restrict__ first, int *__restrict__ flags) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  int32_t e = seg_edge[s];
  if (e < 0) atomicOr(flags, kNotCanonical);
  int32_t prev = s ? seg_edge[s - 1] : -2;
  if (s && e < prev) atomicOr(flags, kNotCanonical);
  first[s] = (s == 0 || e != prev) ? 1 : 0;
}

// run_of[s] is the inclusive scan of first[] (1-based run number); note where runs start
__global__ void run_start_kernel(int64_t nseg, const int32_t *__restrict__ first,
                                 const int32_t *__restrict__ run_of, int32_t *__restrict__ run_start,
                                 int64_t nrun) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s < nseg && first[s]) run_start[run_of[s] - 1] = (int32_t)s;
  if (s == nseg) run_start[nrun] = (int32_t)nseg;
}

__global__ void run_weights_kernel(int64_t nrun, const int32_t *__restrict__ run_start,
                                   longlong2 *__restrict__ wts) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r > nrun) return;
  long long w = r < nrun ? run_start[r + 1] - run_start[r] : 0;
  wts[r] = make_longlong2(w * w, w > 1 ? 1 : 0);
}

struct PairSum {
  __host__ __device__ longlong2 operator()(const longlong2 &a, const longlong2 &b) const {
    return make_longlong2(a.x + b.x, a.y + b.y);
  }
};

__global__ void check_groups_kernel(int64_t ngroup, const int32_t *__restrict__ row,
                                    const int32_t *__restrict__ st, const int32_t *__restrict__ ed,
                                    const int32_t *__restrict__ seg_edge,
                                    const int32_t *__restrict__ run_of,
                                    const int32_t *__restrict__ run_start,
                                    const longlong2 *__restrict__ run_off, int *__restrict__ flags) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= ngroup) return;
  int32_t a = st[g], b = ed[g];
  int32_t run = run_of[a] - 1;
  int32_t base = run_start[run], w = run_start[run + 1] - base;
  bool ok = run_of[b] - 1 == run && row[g] == seg_edge[base] &&
            g == run_off[run].x + (long long)(b - base) * w + (a - base);
  if (!ok) atomicOr(flags, kNotCanonical);
}

__global__ void seg_slot_kernel(int64_t nseg, const int32_t *__restrict__ run_of,
                                const int32_t *__restrict__ run_start,
                                const longlong2 *__restrict__ run_off, int32_t *__restrict__ seg_slot,
                                int32_t *__restrict__ heavy_flag) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  int32_t run = run_of[s] - 1;
  int32_t w = run_start[run + 1] - run_start[run];
  seg_slot[s] = w > 1 ? (int32_t)run_off[run].y : -1;
  heavy_flag[s] = w > 1 ? 1 : 0;
}

__global__ void heavy_list_kernel(int64_t nseg, const int32_t *__restrict__ heavy_flag,
                                  const int32_t *__restrict__ heavy_pos,
                                  int32_t *__restrict__ heavy_segs) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s < nseg && heavy_flag[s]) heavy_segs[heavy_pos[s] - 1] = (int32_t)s;
}

// occurrence count and first position (in H_T_colind order) of every vertex
__global__ void vertex_touch_kernel(int64_t nnz, const int32_t *__restrict__ colind,
                                    int32_t *__restrict__ cnt, int32_t *__restrict__ minpos) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  int32_t v = colind[p];
  atomicAdd(cnt + v, 1);
  atomicMin(minpos + v, (int32_t)p);
}

__global__ void cflag_kernel(int64_t nnz, const int32_t *__restrict__ colind,
                             const int32_t *__restrict__ cnt, const int32_t *__restrict__ minpos,
                             int32_t *__restrict__ cflag) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  int32_t v = colind[p];
  uint32_t c = (uint32_t)v;
  if (minpos[v] == (int32_t)p) c |= 0x80000000u;
  if (cnt[v] == 1) c |= 0x40000000u;
  cflag[p] = (int32_t)c;
}

__global__ void iso_flag_kernel(int64_t n, const int32_t *__restrict__ cnt, int32_t *__restrict__ iso,
                                int32_t *__restrict__ excl) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n) return;
  iso[v] = cnt[v] == 0 ? 1 : 0;
  excl[v] = cnt[v] == 1 ? 1 : 0;
}

__global__ void iso_list_kernel(int64_t n, const int32_t *__restrict__ iso,
                                const int32_t *__restrict__ pos, int32_t *__restrict__ list) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v < n && iso[v]) list[pos[v] - 1] = (int32_t)v;
}

// (vertex, hyperedge) pair of every index position, for the transposition H^T -> H
__global__ void pos_pairs_kernel(int64_t nseg, const int32_t *__restrict__ key, const int32_t *__restrict__ seg_edge,
                                 const int32_t *__restrict__ colind, int64_t *__restrict__ rows,
                                 int64_t *__restrict__ cols) {
  const int lane = threadIdx.x & 31;
  int64_t s = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;   // one warp per segment
  if (s >= nseg) return;
  const int32_t e = seg_edge[s];
  for (int32_t p = key[s] + lane; p < key[s + 1]; p += 32) {
    rows[p] = colind[p];
    cols[p] = e;
  }
}

__global__ void max_deg_kernel(int64_t n, const int32_t *__restrict__ ptr, int *__restrict__ out) {
  int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v < n) atomicMax(out, ptr[v + 1] - ptr[v]);
}

#define GRID(n) (unsigned)ceil_div<int64_t>((n), 256), 256

template <typename In, typename Out, typename Op, typename T>
int exclusive_scan(In in, Out out, Op op, T init, int64_t n, cudaStream_t s) {
  size_t bytes = 0;
  HG_CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, bytes, in, out, op, init, n, s));
  DevBuf<char> ws;
  HG_CUDA_TRY(ws.alloc(bytes));
  HG_CUDA_TRY(cub::DeviceScan::ExclusiveScan(ws.p, bytes, in, out, op, init, n, s));
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  return HG_OK;
}

template <typename In, typename Out>
int inclusive_sum(In in, Out out, int64_t n, cudaStream_t s) {
  size_t bytes = 0;
  HG_CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, bytes, in, out, n, s));
  DevBuf<char> ws;
  HG_CUDA_TRY(ws.alloc(bytes));
  HG_CUDA_TRY(cub::DeviceScan::InclusiveSum(ws.p, bytes, in, out, n, s));
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  return HG_OK;
}

int inclusive_sum_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t s);

// first-touch / single-writer flags for the fused kernel (vertex ids must fit 30 bits)
#ifdef HGEF_LAB
int build_fused(hgPlan *p, cudaStream_t s) {
  const int64_t N = p->num_nodes, Z = p->nnz;
  if (N >= (int64_t(1) << 30) || Z == 0 || N == 0) return HG_OK;  // cflag stays NULL: two-pass path
  DevBuf<int32_t> cnt, minpos, iso, excl, pos;
  HG_CUDA_TRY(cnt.alloc(N)); HG_CUDA_TRY(minpos.alloc(N)); HG_CUDA_TRY(iso.alloc(N));
  HG_CUDA_TRY(excl.alloc(N)); HG_CUDA_TRY(pos.alloc(N));
  HG_CUDA_TRY(cudaMemsetAsync(cnt.p, 0, (size_t)N * sizeof(int32_t), s));
  HG_CUDA_TRY(cudaMemsetAsync(minpos.p, 0x7f, (size_t)N * sizeof(int32_t), s));
  HG_CUDA_TRY(cudaMalloc((void **)&p->cflag, (size_t)Z * sizeof(int32_t)));
  const size_t ctrl_ints = (size_t)p->nseg + (size_t)p->nseg / 32 + 64;  // tiles of >= 1 segment
  HG_CUDA_TRY(cudaMalloc((void **)&p->ctrl, ctrl_ints * sizeof(int32_t)));
  HG_CUDA_TRY(cudaMemsetAsync(p->ctrl, 0, ctrl_ints * sizeof(int32_t), s));
  vertex_touch_kernel<<<GRID(Z), 0, s>>>(Z, p->colind, cnt.p, minpos.p);
  cflag_kernel<<<GRID(Z), 0, s>>>(Z, p->colind, cnt.p, minpos.p, p->cflag);
  iso_flag_kernel<<<GRID(N), 0, s>>>(N, cnt.p, iso.p, excl.p);
  HG_CUDA_TRY(cudaGetLastError());
  int32_t niso = 0, nexcl = 0;
  if (int rc = inclusive_sum_i32(excl.p, pos.p, N, s)) return rc;
  HG_CUDA_TRY(cudaMemcpy(&nexcl, pos.p + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (int rc = inclusive_sum_i32(iso.p, pos.p, N, s)) return rc;
  HG_CUDA_TRY(cudaMemcpy(&niso, pos.p + (N - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
  p->niso = niso; p->nexcl = nexcl;
  if (niso) {
    HG_CUDA_TRY(cudaMalloc((void **)&p->iso_list, (size_t)niso * sizeof(int32_t)));
    iso_list_kernel<<<GRID(N), 0, s>>>(N, iso.p, pos.p, p->iso_list);
    HG_CUDA_TRY(cudaGetLastError());
  }
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  return HG_OK;
}

#endif  // HGEF_LAB

// H (vertex -> hyperedges, ascending) by transposing the caller's H^T with the CSR builder
int build_pull(hgPlan *p, cudaStream_t s) {
  const int64_t N = p->num_nodes, M = p->num_edges, Z = p->nnz;
  if (Z == 0 || N == 0 || M == 0) return HG_OK;
  DevBuf<int64_t> rows, cols;
  DevBuf<int32_t> t_ptr, t_ind;
  DevBuf<float> data, t_data;
  DevBuf<int> mx;
  HG_CUDA_TRY(rows.alloc(Z)); HG_CUDA_TRY(cols.alloc(Z));
  HG_CUDA_TRY(t_ptr.alloc(M + 1)); HG_CUDA_TRY(t_ind.alloc(Z));
  HG_CUDA_TRY(data.alloc(Z)); HG_CUDA_TRY(t_data.alloc(Z)); HG_CUDA_TRY(mx.alloc(1));
  HG_CUDA_TRY(cudaMalloc((void **)&p->h_ptr, (size_t)(N + 1) * sizeof(int32_t)));
  HG_CUDA_TRY(cudaMalloc((void **)&p->h_ind, (size_t)Z * sizeof(int32_t)));
  pos_pairs_kernel<<<(unsigned)ceil_div<int64_t>(p->nseg * 32, 256), 256, 0, s>>>(p->nseg, p->key, p->seg_edge,
                                                                                  p->colind, rows.p, cols.p);
  HG_CUDA_TRY(cudaGetLastError());
  int64_t z = 0;
  if (int rc = hg_csr_build_dev(N, M, Z, rows.p, cols.p, p->h_ptr, p->h_ind, data.p, t_ptr.p, t_ind.p, t_data.p,
                                &z, p->device, (void *)s))
    return rc;
  if (z != Z) {  // a vertex listed twice in one hyperedge: the pull form would count it once
    cudaFree(p->h_ptr); cudaFree(p->h_ind);
    p->h_ptr = p->h_ind = nullptr;
    return HG_OK;
  }
  HG_CUDA_TRY(cudaMemsetAsync(mx.p, 0, sizeof(int), s));
  max_deg_kernel<<<GRID(N), 0, s>>>(N, p->h_ptr, mx.p);
  HG_CUDA_TRY(cudaMemcpyAsync(&p->max_vdeg, mx.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  return HG_OK;
}

int build(hgPlan *p, cudaStream_t s) {
  const int64_t S = p->nseg, G = p->ngroup;
  DevBuf<int> flags, max_len;
  HG_CUDA_TRY(flags.alloc(1));
  HG_CUDA_TRY(max_len.alloc(1));
  HG_CUDA_TRY(cudaMemsetAsync(flags.p, 0, sizeof(int), s));
  HG_CUDA_TRY(cudaMemsetAsync(max_len.p, 0, sizeof(int), s));
  HG_CUDA_TRY(cudaMalloc((void **)&p->seg_edge, (size_t)S * sizeof(int32_t)));
  HG_CUDA_TRY(cudaMalloc((void **)&p->seg_slot, (size_t)S * sizeof(int32_t)));
  HG_CUDA_TRY(cudaMemsetAsync(p->seg_edge, 0xff, (size_t)S * sizeof(int32_t), s));

  if (p->nnz) check_colind_kernel<<<GRID(p->nnz), 0, s>>>(p->nnz, p->colind, p->num_nodes, flags.p);
  check_key_kernel<<<GRID(S), 0, s>>>(S, p->key, p->nnz, flags.p, max_len.p);
  if (G) diag_groups_kernel<<<GRID(G), 0, s>>>(G, p->row, p->st, p->ed, S, p->num_edges, p->seg_edge,
                                               flags.p);
  HG_CUDA_TRY(cudaGetLastError());
  int h_flags = 0;
  HG_CUDA_TRY(cudaMemcpyAsync(&h_flags, flags.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  HG_CUDA_TRY(cudaMemcpyAsync(&p->max_seg_len, max_len.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  HG_CUDA_TRY(cudaStreamSynchronize(s));
  if (h_flags & kBadIndex)
    return set_error(HG_EGRAPH, "plan: H_T_colind holds a vertex id outside [0, %lld)",
                     (long long)p->num_nodes);
  if (h_flags & kBadKey)
    return set_error(HG_EGRAPH, "plan: group_key is not a non-decreasing offset array within [0, %lld]",
                     (long long)p->nnz);
  if (h_flags & kBadGroup)
    return set_error(HG_EGRAPH, "plan: a group names a segment outside [0, %lld) or a hyperedge "
                                "outside [0, %lld)", (long long)S, (long long)p->num_edges);

  DevBuf<int32_t> first, run_of, run_start, heavy_flag, heavy_pos;
  HG_CUDA_TRY(first.alloc(S));
  HG_CUDA_TRY(run_of.alloc(S));
  seg_first_kernel<<<GRID(S), 0, s>>>(S, p->seg_edge, first.p, flags.p);
  HG_CUDA_TRY(cudaGetLastError());
  if (int rc = inclusive_sum(first.p, run_of.p, S, s)) return rc;
  int32_t nrun = 0;
  HG_CUDA_TRY(cudaMemcpy(&nrun, run_of.p + (S - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
  HG_CUDA_TRY(cudaMemcpy(&h_flags, flags.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (h_flags & kNotCanonical) {  // a segment without a diagonal group: literal schedule only
    p->canonical = 0;
    return HG_OK;
  }
  HG_CUDA_TRY(run_start.alloc((size_t)nrun + 1));
  run_start_kernel<<<GRID(S + 1), 0, s>>>(S, first.p, run_of.p, run_start.p, nrun);
  DevBuf<longlong2> wts, run_off;
  HG_CUDA_TRY(wts.alloc((size_t)nrun + 1));
  HG_CUDA_TRY(run_off.alloc((size_t)nrun + 1));
  run_weights_kernel<<<GRID(nrun + 1), 0, s>>>(nrun, run_start.p, wts.p);
  HG_CUDA_TRY(cudaGetLastError());
  if (int rc = exclusive_scan(wts.p, run_off.p, PairSum(), make_longlong2(0, 0), (int64_t)nrun + 1, s))
    return rc;
  longlong2 tot;
  HG_CUDA_TRY(cudaMemcpy(&tot, run_off.p + nrun, sizeof(tot), cudaMemcpyDeviceToHost));
  if (tot.x != G) {
    p->canonical = 0;
    return HG_OK;
  }
  if (G) check_groups_kernel<<<GRID(G), 0, s>>>(G, p->row, p->st, p->ed, p->seg_edge, run_of.p,
                                                run_start.p, run_off.p, flags.p);
  HG_CUDA_TRY(heavy_flag.alloc(S));
  HG_CUDA_TRY(heavy_pos.alloc(S));
  seg_slot_kernel<<<GRID(S), 0, s>>>(S, run_of.p, run_start.p, run_off.p, p->seg_slot, heavy_flag.p);
  HG_CUDA_TRY(cudaGetLastError());
  if (int rc = inclusive_sum(heavy_flag.p, heavy_pos.p, S, s)) return rc;
  int32_t nheavy_segs = 0;
  HG_CUDA_TRY(cudaMemcpy(&nheavy_segs, heavy_pos.p + (S - 1), sizeof(int32_t), cudaMemcpyDeviceToHost));
  HG_CUDA_TRY(cudaMemcpy(&h_flags, flags.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (h_flags & kNotCanonical) {
    p->canonical = 0;
    return HG_OK;
  }
  p->canonical = 1;
  p->nheavy_edges = tot.y;
  p->nheavy_segs = nheavy_segs;
  if (nheavy_segs) {
    HG_CUDA_TRY(cudaMalloc((void **)&p->heavy_segs, (size_t)nheavy_segs * sizeof(int32_t)));
    heavy_list_kernel<<<GRID(S), 0, s>>>(S, heavy_flag.p, heavy_pos.p, p->heavy_segs);
    HG_CUDA_TRY(cudaGetLastError());
    HG_CUDA_TRY(cudaStreamSynchronize(s));
  }
  // the optional forms need segments that tile [0, nnz) exactly (the balancer's do; caller-built canonical
  // group arrays need not): otherwise only the two-pass form is available
  {
    int32_t k0 = -1, kS = -1;
    HG_CUDA_TRY(cudaMemcpy(&k0, p->key, sizeof(int32_t), cudaMemcpyDeviceToHost));
    HG_CUDA_TRY(cudaMemcpy(&kS, p->key + S, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (k0 != 0 || kS != p->nnz) return HG_OK;
  }
#ifdef HGEF_LAB
  if (int rc = build_fused(p, s)) return rc;
#endif
  if (int rc = build_pull(p, s)) return rc;
  return build_stream(p, s);
}

int inclusive_sum_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t s) {
  return inclusive_sum(in, out, n, s);
}

}  // namespace
}  // namespace hg

using namespace hg;

namespace hg {
int plan_grow(hgPlan *p, float **buf, size_t *cap, size_t need, cudaStream_t s, const char *what) {
  if (need <= *cap) return HG_OK;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone)
    return set_error(HG_EINVAL, "the plan's %s buffer must grow to %zu bytes, which cannot happen while the stream is "
                     "being captured: call hg_plan_reserve (or run the op once eagerly) first", what, need * sizeof(float));
  float *nb = nullptr;
  if (cudaMalloc((void **)&nb, need * sizeof(float)) != cudaSuccess) {
    cudaGetLastError();
    return set_error(HG_ENOMEM, "cannot allocate %zu bytes for the plan's %s", need * sizeof(float), what);
  }
  if (*buf) p->retired.push_back(*buf);
  *buf = nb;
  *cap = need;
  return HG_OK;
}
}  // namespace hg

extern "C" {

int hg_plan_create(hgPlan **plan, int64_t num_nodes, int64_t num_edges, int64_t nnz, int64_t nseg,
                   int64_t ngroup, const int32_t *d_key, const int32_t *d_row, const int32_t *d_st,
                   const int32_t *d_ed, const int32_t *d_t_indices, int device, void *stream) {
  HG_REQUIRE(plan != nullptr, "plan_create: plan is NULL");
  *plan = nullptr;
  HG_REQUIRE(num_nodes >= 0 && num_edges >= 0 && nnz >= 0 && nseg >= 1 && ngroup >= 0,
             "plan_create: negative size (or no segment)");
  HG_REQUIRE(num_nodes <= INT32_MAX && nnz <= INT32_MAX && nseg < INT32_MAX && ngroup <= INT32_MAX,
             "plan_create: sizes exceed int32 index arrays");
  HG_REQUIRE(d_key && (d_t_indices || nnz == 0), "plan_create: group_key / H_T_colind is NULL");
  HG_REQUIRE((d_row && d_st && d_ed) || ngroup == 0, "plan_create: a group array is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "plan_create: cannot select device %d", device);
  hgPlan *p = new (std::nothrow) hgPlan();
  if (!p) return set_error(HG_ENOMEM, "plan_create: out of host memory");
  HG_CUDA_TRY(cudaGetDevice(&p->device));
  cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, p->device);
  p->num_nodes = num_nodes; p->num_edges = num_edges; p->nnz = nnz; p->nseg = nseg; p->ngroup = ngroup;
  p->key = d_key; p->row = d_row; p->st = d_st; p->ed = d_ed; p->colind = d_t_indices;
  int rc = build(p, (cudaStream_t)stream);
  if (rc != HG_OK) {
    hg_plan_destroy(p);
    return rc;
  }
  *plan = p;
  return HG_OK;
}

int hg_plan_destroy(hgPlan *p) {
  if (!p) return HG_OK;
  DeviceGuard guard(p->device);
  cudaFree(p->seg_edge);
  cudaFree(p->seg_slot);
  cudaFree(p->heavy_segs);
  cudaFree(p->h_ptr);
  cudaFree(p->h_ind);
  cudaFree(p->xe);
  cudaFree(p->cflag);
  cudaFree(p->iso_list);
  cudaFree(p->ctrl);
  cudaFree(p->scratch);
  cudaFree(p->pad_x);
  cudaFree(p->pad_y);
  for (void *q : p->retired) cudaFree(q);
  stream_free(p);
#ifdef HGEF_LAB
  ring_free(p);
#endif
  delete p;
  return HG_OK;
}

int hg_plan_info(const hgPlan *p, int64_t *nseg, int64_t *nheavy_edges, int64_t *nheavy_segs,
                 int32_t *canonical) {
  HG_REQUIRE(p != nullptr, "plan_info: plan is NULL");
  if (nseg) *nseg = p->nseg;
  if (nheavy_edges) *nheavy_edges = p->nheavy_edges;
  if (nheavy_segs) *nheavy_segs = p->nheavy_segs;
  if (canonical) *canonical = p->canonical;
  return HG_OK;
}

}  // extern "C"
