// hgef_csr.cu -- incidence pairs -> CSR of H and of H^T, and the degree scalings.
//
// Contract: the arrays scipy produces for HyperGsys/hypergraph.py:23-25
//     H   = coo_matrix((ones, (V, E)), (N, M)).tocsr()
//     H_T = H.transpose().tocsr()
// i.e. column indices ascending inside each row, duplicate (row, col) pairs merged into
// one entry whose value is the multiplicity, int32 indptr/indices.
//
// Design: a (row, col) pair is one 64-bit key row<<32|col.  Sorting the keys and
// run-length encoding them gives H directly; swapping the halves of the unique keys and
// sorting again gives H^T.  On the device both sorts are CUB radix sorts limited to the
// significant bits; indptr is filled from the row boundaries of the sorted keys.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>

#include <algorithm>
#include <climits>
#include <vector>

#include "hgef_common.cuh"

namespace hg {
namespace {

int check_shape(int64_t nrow, int64_t ncol, int64_t nnz) {
  HG_REQUIRE(nrow >= 0 && ncol >= 0 && nnz >= 0, "csr_build: negative size");
  HG_REQUIRE(nrow < INT32_MAX && ncol < INT32_MAX && nnz <= INT32_MAX,
             "csr_build: %lld x %lld with %lld non-zeros does not fit int32 CSR arrays",
             (long long)nrow, (long long)ncol, (long long)nnz);
  return HG_OK;
}

inline int bits_for(int64_t n) {  // bits needed to represent values in [0, n)
  int b = 1;
  while (b < 32 && (int64_t(1) << b) < n) ++b;
  return b;
}

// ---------------------------------------------------------------- host
// sorted unique keys + multiplicities -> one CSR
void emit_csr(const std::vector<uint64_t> &keys, const std::vector<float> &val, int64_t nrow,
              int32_t *indptr, int32_t *indices, float *data) {
  const int64_t n = (int64_t)keys.size();
  int64_t p = 0;
  for (int64_t r = 0; r <= nrow; ++r) {
    while (p < n && (int64_t)(keys[p] >> 32) < r) ++p;
    indptr[r] = (int32_t)p;
  }
  for (int64_t q = 0; q < n; ++q) {
    indices[q] = (int32_t)(keys[q] & 0xffffffffu);
    data[q] = val[q];
  }
}

// ---------------------------------------------------------------- device
__global__ void pack_keys_kernel(int64_t n, const int64_t *__restrict__ rows,
                                 const int64_t *__restrict__ cols, int64_t nrow, int64_t ncol,
                                 uint64_t *__restrict__ keys, int *__restrict__ bad) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  int64_t r = rows[p], c = cols[p];
  if (r < 0 || r >= nrow || c < 0 || c >= ncol) {
    *bad = 1;
    r = 0; c = 0;
  }
  keys[p] = ((uint64_t)r << 32) | (uint64_t)c;
}

__global__ void swap_halves_kernel(int64_t n, const uint64_t *__restrict__ in,
                                   const int32_t *__restrict__ cnt, uint64_t *__restrict__ out,
                                   float *__restrict__ val) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint64_t k = in[p];
  out[p] = (k << 32) | (k >> 32);
  val[p] = (float)cnt[p];
}

// keys sorted by (major, minor): indices = minor, indptr[r] = first position whose major >= r
__global__ void unpack_kernel(int64_t n, int64_t nmajor, const uint64_t *__restrict__ keys,
                              int32_t *__restrict__ indptr, int32_t *__restrict__ indices) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p > n) return;
  int64_t cur = p < n ? (int64_t)(keys[p] >> 32) : nmajor;
  int64_t prev = p > 0 ? (int64_t)(keys[p - 1] >> 32) : -1;
  for (int64_t r = prev + 1; r <= cur; ++r) indptr[r] = (int32_t)p;
  if (p < n) indices[p] = (int32_t)(keys[p] & 0xffffffffu);
}

__global__ void cnt_to_float_kernel(int64_t n, const int32_t *__restrict__ cnt, float *__restrict__ v) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < n) v[p] = (float)cnt[p];
}

__global__ void degree_scale_kernel(int64_t nrow, const int32_t *__restrict__ indptr,
                                    const float *__restrict__ data, float power, int inf_to_one,
                                    float *__restrict__ out) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= nrow) return;
  // scipy sums float64 ones (hypergraph.py:34-35) and torch casts to float (:36-38):
  // the degree is an exact small integer either way.
  double d = 0;
  if (data) {
    for (int32_t p = indptr[r]; p < indptr[r + 1]; ++p) d += (double)data[p];
  } else {
    d = (double)(indptr[r + 1] - indptr[r]);
  }
  float x = (float)d, y;
  // torch.pow(float, -0.5 / -1) on the host: correctly rounded reciprocal (sqrt)
  if (power == -0.5f) y = 1.0f / sqrtf(x);
  else if (power == -1.0f) y = 1.0f / x;
  else y = powf(x, power);
  if (inf_to_one && isinf(y)) y = 1.0f;
  out[r] = y;
}

#define GRID(n) (unsigned)ceil_div<int64_t>((n), 256), 256

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

int hg_csr_build_host(int64_t nrow, int64_t ncol, int64_t nnz_in, const int64_t *h_rows,
                      const int64_t *h_cols, int32_t *h_indptr, int32_t *h_indices, float *h_data,
                      int32_t *h_t_indptr, int32_t *h_t_indices, float *h_t_data,
                      int64_t *nnz_out) {
  if (int rc = check_shape(nrow, ncol, nnz_in)) return rc;
  HG_REQUIRE((h_rows && h_cols) || nnz_in == 0, "csr_build: rows/cols are NULL");
  HG_REQUIRE(h_indptr && h_t_indptr && nnz_out, "csr_build: an output pointer is NULL");
  HG_REQUIRE((h_indices && h_data && h_t_indices && h_t_data) || nnz_in == 0,
             "csr_build: an output array is NULL");
  std::vector<uint64_t> keys((size_t)nnz_in);
  for (int64_t p = 0; p < nnz_in; ++p) {
    const int64_t r = h_rows[p], c = h_cols[p];
    HG_REQUIRE(r >= 0 && r < nrow && c >= 0 && c < ncol,
               "csr_build: pair %lld = (%lld, %lld) outside %lld x %lld", (long long)p,
               (long long)r, (long long)c, (long long)nrow, (long long)ncol);
    keys[(size_t)p] = ((uint64_t)r << 32) | (uint64_t)c;
  }
  std::sort(keys.begin(), keys.end());
  std::vector<uint64_t> uniq;
  std::vector<float> mult;
  uniq.reserve(keys.size());
  mult.reserve(keys.size());
  for (size_t p = 0; p < keys.size(); ++p) {
    if (!uniq.empty() && uniq.back() == keys[p]) mult.back() += 1.0f;
    else { uniq.push_back(keys[p]); mult.push_back(1.0f); }
  }
  emit_csr(uniq, mult, nrow, h_indptr, h_indices, h_data);
  // transpose: swap halves, sort (stable pairing of the multiplicities through an index sort)
  const size_t z = uniq.size();
  std::vector<uint64_t> tk(z);
  std::vector<uint32_t> order(z);
  for (size_t p = 0; p < z; ++p) { tk[p] = (uniq[p] << 32) | (uniq[p] >> 32); order[p] = (uint32_t)p; }
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return tk[a] < tk[b]; });
  std::vector<uint64_t> tks(z);
  std::vector<float> tv(z);
  for (size_t p = 0; p < z; ++p) { tks[p] = tk[order[p]]; tv[p] = mult[order[p]]; }
  emit_csr(tks, tv, ncol, h_t_indptr, h_t_indices, h_t_data);
  *nnz_out = (int64_t)z;
  return HG_OK;
}

int hg_csr_build_dev(int64_t nrow, int64_t ncol, int64_t nnz_in, const int64_t *d_rows,
                     const int64_t *d_cols, int32_t *d_indptr, int32_t *d_indices, float *d_data,
                     int32_t *d_t_indptr, int32_t *d_t_indices, float *d_t_data, int64_t *nnz_out,
                     int device, void *stream) {
  if (int rc = check_shape(nrow, ncol, nnz_in)) return rc;
  HG_REQUIRE((d_rows && d_cols) || nnz_in == 0, "csr_build: rows/cols are NULL");
  HG_REQUIRE(d_indptr && d_t_indptr && nnz_out, "csr_build: an output pointer is NULL");
  HG_REQUIRE((d_indices && d_data && d_t_indices && d_t_data) || nnz_in == 0,
             "csr_build: an output array is NULL");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "csr_build: cannot select device %d", device);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = nnz_in;
  DevBuf<uint64_t> k0, k1, uniq;
  DevBuf<int32_t> cnt, nruns;
  DevBuf<float> tv0;
  DevBuf<int> bad;
  DevBuf<char> ws;
  HG_CUDA_TRY(k0.alloc(n)); HG_CUDA_TRY(k1.alloc(n)); HG_CUDA_TRY(uniq.alloc(n));
  HG_CUDA_TRY(cnt.alloc(n)); HG_CUDA_TRY(nruns.alloc(1)); HG_CUDA_TRY(tv0.alloc(n));
  HG_CUDA_TRY(bad.alloc(1));
  HG_CUDA_TRY(cudaMemsetAsync(bad.p, 0, sizeof(int), s));
  HG_CUDA_TRY(cudaMemsetAsync(nruns.p, 0, sizeof(int32_t), s));
  int64_t z = 0;
  if (n > 0) {
    pack_keys_kernel<<<GRID(n), 0, s>>>(n, d_rows, d_cols, nrow, ncol, k0.p, bad.p);
    HG_CUDA_TRY(cudaGetLastError());
    const int end_bit = 32 + bits_for(nrow);
    size_t b_sort = 0, b_rle = 0, b_pair = 0;
    HG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(nullptr, b_sort, k0.p, k1.p, n, 0, end_bit, s));
    HG_CUDA_TRY(cub::DeviceRunLengthEncode::Encode(nullptr, b_rle, k1.p, uniq.p, cnt.p, nruns.p, n, s));
    const int end_bit_t = 32 + bits_for(ncol);
    HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, b_pair, k0.p, k1.p, tv0.p, d_t_data, n, 0,
                                                end_bit_t, s));
    HG_CUDA_TRY(ws.alloc(std::max(b_sort, std::max(b_rle, b_pair))));
    HG_CUDA_TRY(cub::DeviceRadixSort::SortKeys(ws.p, b_sort, k0.p, k1.p, n, 0, end_bit, s));
    HG_CUDA_TRY(cub::DeviceRunLengthEncode::Encode(ws.p, b_rle, k1.p, uniq.p, cnt.p, nruns.p, n, s));
    int h_bad = 0;
    int32_t h_runs = 0;
    HG_CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    HG_CUDA_TRY(cudaMemcpyAsync(&h_runs, nruns.p, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    HG_CUDA_TRY(cudaStreamSynchronize(s));
    HG_REQUIRE(!h_bad, "csr_build: an incidence pair lies outside %lld x %lld", (long long)nrow,
               (long long)ncol);
    z = h_runs;
    // H
    unpack_kernel<<<GRID(z + 1), 0, s>>>(z, nrow, uniq.p, d_indptr, d_indices);
    cnt_to_float_kernel<<<GRID(z), 0, s>>>(z, cnt.p, d_data);
    // H^T
    swap_halves_kernel<<<GRID(z), 0, s>>>(z, uniq.p, cnt.p, k0.p, tv0.p);
    HG_CUDA_TRY(cudaGetLastError());
    HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws.p, b_pair, k0.p, k1.p, tv0.p, d_t_data, z, 0,
                                                end_bit_t, s));
    unpack_kernel<<<GRID(z + 1), 0, s>>>(z, ncol, k1.p, d_t_indptr, d_t_indices);
    HG_CUDA_TRY(cudaGetLastError());
  } else {
    HG_CUDA_TRY(cudaMemsetAsync(d_indptr, 0, (size_t)(nrow + 1) * sizeof(int32_t), s));
    HG_CUDA_TRY(cudaMemsetAsync(d_t_indptr, 0, (size_t)(ncol + 1) * sizeof(int32_t), s));
  }
  HG_CUDA_TRY(cudaStreamSynchronize(s));  // scratch is freed on return
  *nnz_out = z;
  return HG_OK;
}

int hg_degree_scale_dev(int64_t nrow, const int32_t *d_indptr, const float *d_data, float power,
                        int inf_to_one, float *d_out, int device, void *stream) {
  HG_REQUIRE(nrow >= 0 && d_indptr && d_out, "degree_scale: bad arguments");
  DeviceGuard guard(device);
  HG_REQUIRE(guard.ok(), "degree_scale: cannot select device %d", device);
  if (nrow == 0) return HG_OK;
  degree_scale_kernel<<<GRID(nrow), 0, (cudaStream_t)stream>>>(nrow, d_indptr, d_data, power,
                                                               inf_to_one, d_out);
  HG_CUDA_TRY(cudaGetLastError());
  return HG_OK;
}

}  // extern "C"
