// hgef_balancer.cu -- the workload balancer, host and device.
//
// Contract (bit-exact with HyperGsys/balancer.py:15-33 and its C++ twin
// include/taskbalancer/balancer_kernel.cuh:229-259): every hyperedge row of H^T with
// deg non-zeros becomes w = ceil(deg/ngs) segments of ngs entries (the last one shorter);
// `key` lists the segment start offsets followed by the sentinel csrptr[nrow]; and for
// every ordered pair (write segment i, read segment j) of one row there is one group
// (st = base+j, ed = base+i, row = rid), pairs in i-major order.
//
// The reference builds these with an O(G) interpreted loop.  Both versions here are
// closed-form: the position of every output element is a function of two prefix sums
// (sum of w, sum of w^2), so the device version is one scan plus two embarrassingly
// parallel fills, balanced over OUTPUT elements (a giant hyperedge with w^2 groups does
// not serialise on one thread).
#include <cub/device/device_scan.cuh>

#include <climits>
#include <vector>

#include "hgef_common.cuh"

namespace hg {
namespace {

inline int64_t seg_count(int64_t lo, int64_t hi, int64_t ngs) { return (hi - lo + ngs - 1) / ngs; }

int check_args(int64_t nrow, const void *csrptr, int32_t ngs) {
  HG_REQUIRE(nrow >= 0, "balancer: nrow must be >= 0 (got %lld)", (long long)nrow);
  HG_REQUIRE(csrptr != nullptr, "balancer: csrptr is NULL");
  HG_REQUIRE(ngs >= 1, "balancer: ngs must be >= 1 (got %d)", ngs);
  return HG_OK;
}

// ---------------------------------------------------------------- device
struct SegOp {  // per-row (w, w*w) packed for a single scan
  const int32_t *csrptr;
  int32_t ngs;
  __host__ __device__ longlong2 operator()(int64_t r) const {
    long long w = ((long long)csrptr[r + 1] - csrptr[r] + ngs - 1) / ngs;
    if (w < 0) w = 0;
    return make_longlong2(w, w * w);
  }
};
struct PairSum {
  __host__ __device__ longlong2 operator()(const longlong2 &a, const longlong2 &b) const {
    return make_longlong2(a.x + b.x, a.y + b.y);
  }
};

__global__ void row_weights_kernel(int64_t nrow, SegOp op, longlong2 *out) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < nrow) out[r] = op(r);
  if (r == nrow) out[r] = make_longlong2(0, 0);  // slot for the grand total
}

// first row index whose exclusive prefix exceeds x, minus one; rows with w == 0 share
// their offset with the next row and are skipped by construction.
template <bool kGroups>
__device__ __forceinline__ int64_t find_row(const longlong2 *off, int64_t nrow, long long x) {
  int64_t lo = 0, hi = nrow;  // off[nrow] = total > x
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    long long v = kGroups ? off[mid + 1].y : off[mid + 1].x;
    if (v > x) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__global__ void fill_keys_kernel(int64_t nrow, const int32_t *__restrict__ csrptr, int32_t ngs,
                                 const longlong2 *__restrict__ off, int64_t nseg,
                                 int32_t *__restrict__ key) {
  int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s < nseg) {
    int64_t r = find_row<false>(off, nrow, s);
    key[s] = (int32_t)(csrptr[r] + (s - off[r].x) * ngs);
  } else if (s == nseg) {
    key[s] = csrptr[nrow];
  }
}

__global__ void fill_groups_kernel(int64_t nrow, const longlong2 *__restrict__ off, int64_t ngroup,
                                   int32_t *__restrict__ row, int32_t *__restrict__ st,
                                   int32_t *__restrict__ ed) {
  int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= ngroup) return;
  int64_t r = find_row<true>(off, nrow, g);
  long long base = off[r].x, w = off[r + 1].x - base, local = g - off[r].y;
  long long i = local / w, j = local - i * w;
  row[g] = (int32_t)r;
  st[g] = (int32_t)(base + j);
  ed[g] = (int32_t)(base + i);
}

// exclusive prefix (sum w, sum w^2) over rows into off[0..nrow]; off[nrow] = totals.
int scan_rows(int64_t nrow, const int32_t *d_csrptr, int32_t ngs, DevBuf<longlong2> &off,
              cudaStream_t stream) {
  DevBuf<longlong2> tmp;
  HG_CUDA_TRY(off.alloc(nrow + 1));
  HG_CUDA_TRY(tmp.alloc(nrow + 1));
  SegOp op{d_csrptr, ngs};
  row_weights_kernel<<<(unsigned)ceil_div<int64_t>(nrow + 1, 256), 256, 0, stream>>>(nrow, op, tmp.p);
  HG_CUDA_TRY(cudaGetLastError());
  size_t bytes = 0;
  HG_CUDA_TRY(cub::DeviceScan::ExclusiveScan(nullptr, bytes, tmp.p, off.p, PairSum(),
                                             make_longlong2(0, 0), nrow + 1, stream));
  DevBuf<char> ws;
  HG_CUDA_TRY(ws.alloc(bytes));
  HG_CUDA_TRY(cub::DeviceScan::ExclusiveScan(ws.p, bytes, tmp.p, off.p, PairSum(),
                                             make_longlong2(0, 0), nrow + 1, stream));
  HG_CUDA_TRY(cudaStreamSynchronize(stream));  // scratch is freed on return
  return HG_OK;
}

int size_checks(int64_t s, int64_t g) {
  if (s == 0)
    return set_error(HG_EEMPTY, "balancer: no row has a non-zero (the reference raises IndexError, "
                                "balancer.py:32)");
  HG_REQUIRE(s + 1 <= INT32_MAX && g <= INT32_MAX,
             "balancer: %lld segments / %lld groups do not fit int32 index arrays",
             (long long)s, (long long)g);
  return HG_OK;
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

int hg_balance_count_host(int64_t nrow, const int32_t *h_csrptr, int32_t ngs, int64_t *nkey,
                          int64_t *ngroup) {
  if (int rc = check_args(nrow, h_csrptr, ngs)) return rc;
  HG_REQUIRE(nkey && ngroup, "balancer: output size pointers are NULL");
  int64_t s = 0, g = 0;
  for (int64_t r = 0; r < nrow; ++r) {
    HG_REQUIRE(h_csrptr[r + 1] >= h_csrptr[r], "balancer: csrptr decreases at row %lld", (long long)r);
    int64_t w = seg_count(h_csrptr[r], h_csrptr[r + 1], ngs);
    s += w;
    g += w * w;
  }
  if (int rc = size_checks(s, g)) return rc;
  *nkey = s + 1;
  *ngroup = g;
  return HG_OK;
}

int hg_balance_fill_host(int64_t nrow, const int32_t *h_csrptr, int32_t ngs, int32_t *h_key,
                         int32_t *h_row, int32_t *h_st, int32_t *h_ed) {
  if (int rc = check_args(nrow, h_csrptr, ngs)) return rc;
  HG_REQUIRE(h_key && h_row && h_st && h_ed, "balancer: an output array is NULL");
  int32_t *kp = h_key;
  int64_t gp = 0;
  int32_t base = 0;
  for (int64_t r = 0; r < nrow; ++r) {
    const int32_t lo = h_csrptr[r], hi = h_csrptr[r + 1];
    const int32_t w = (int32_t)seg_count(lo, hi, ngs);
    for (int32_t t = 0; t < w; ++t) *kp++ = lo + t * ngs;
    if (w == 1) {  // the overwhelmingly common row
      h_row[gp] = (int32_t)r; h_st[gp] = base; h_ed[gp] = base; ++gp;
    } else {
      for (int32_t i = 0; i < w; ++i)
        for (int32_t j = 0; j < w; ++j, ++gp) {
          h_row[gp] = (int32_t)r; h_st[gp] = base + j; h_ed[gp] = base + i;
        }
    }
    base += w;
  }
  if (kp == h_key)
    return set_error(HG_EEMPTY, "balancer: no row has a non-zero (the reference raises IndexError, "
                                "balancer.py:32)");
  *kp = h_csrptr[nrow];
  return HG_OK;
}

int hg_balance_count_dev(int64_t nrow, const int32_t *d_csrptr, int32_t ngs, int64_t *nkey,
                         int64_t *ngroup, int device, void *stream) {
  if (int rc = check_args(nrow, d_csrptr, ngs)) return rc;
  HG_REQUIRE(nkey && ngroup, "balancer: output size pointers are NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "balancer: cannot select device %d", device);
  DevBuf<longlong2> off;
  if (int rc = scan_rows(nrow, d_csrptr, ngs, off, (cudaStream_t)stream)) return rc;
  longlong2 tot;
  HG_CUDA_TRY(cudaMemcpy(&tot, off.p + nrow, sizeof(tot), cudaMemcpyDeviceToHost));
  if (int rc = size_checks(tot.x, tot.y)) return rc;
  *nkey = tot.x + 1;
  *ngroup = tot.y;
  return HG_OK;
}

int hg_balance_fill_dev(int64_t nrow, const int32_t *d_csrptr, int32_t ngs, int64_t nkey,
                        int64_t ngroup, int32_t *d_key, int32_t *d_row, int32_t *d_st,
                        int32_t *d_ed, int device, void *stream) {
  if (int rc = check_args(nrow, d_csrptr, ngs)) return rc;
  HG_REQUIRE(d_key && d_row && d_st && d_ed, "balancer: an output array is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "balancer: cannot select device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  DevBuf<longlong2> off;
  if (int rc = scan_rows(nrow, d_csrptr, ngs, off, s)) return rc;
  longlong2 tot;
  HG_CUDA_TRY(cudaMemcpy(&tot, off.p + nrow, sizeof(tot), cudaMemcpyDeviceToHost));
  HG_REQUIRE(tot.x + 1 == nkey && tot.y == ngroup,
             "balancer: sizes (%lld, %lld) do not match the count pass (%lld, %lld)",
             (long long)nkey, (long long)ngroup, (long long)tot.x + 1, (long long)tot.y);
  const int64_t nseg = nkey - 1;
  fill_keys_kernel<<<(unsigned)ceil_div<int64_t>(nseg + 1, 256), 256, 0, s>>>(nrow, d_csrptr, ngs,
                                                                             off.p, nseg, d_key);
  HG_CUDA_TRY(cudaGetLastError());
  if (ngroup > 0) {
    fill_groups_kernel<<<(unsigned)ceil_div<int64_t>(ngroup, 256), 256, 0, s>>>(nrow, off.p, ngroup,
                                                                               d_row, d_st, d_ed);
    HG_CUDA_TRY(cudaGetLastError());
  }
  HG_CUDA_TRY(cudaStreamSynchronize(s));  // `off` is freed on return
  return HG_OK;
}

}  // extern "C"
