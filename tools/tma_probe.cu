// tma_probe.cu -- what does it cost to gather feature rows with cp.async.bulk (UBLKCP), and how does the
// throughput depend on WHO issues the copies?  (Design input for the ring form, hgef_ring.cu.)
//
// Every warp streams random rows of a 1.5 GB matrix through a private shared-memory ring: S stages of R rows,
// one mbarrier per stage, copies of stage c + S - 1 issued before stage c is summed from shared memory.
// Variants: the copies of a stage are issued by ONE elected lane (a loop of R UBLKCP) or by R lanes at once
// (one UBLKCP per lane, which ptxas serialises with an ELECT loop); 1..16 such warps per CTA; 1..2 CTAs per SM.
// Reports GB/s and the average clocks the issuing lane spends per UBLKCP.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_row(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t e;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(e));
  return e != 0;
}

struct P {
  const float *X; const int *idx; float *out; unsigned long long *clk;
  int R, S, row_bytes, chunks, lanes_issue;   // lanes_issue: 1 = elected lane loops, else every lane < R issues one
};

__global__ void __launch_bounds__(512) probe(P p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint32_t stage_bytes = (uint32_t)p.R * p.row_bytes;
  unsigned char *ring = smem + 4096 + (size_t)warp * p.S * stage_bytes;
  int *stage_idx = reinterpret_cast<int *>(smem + 1024) + warp * 32;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem) + warp * 8;   // S <= 8
  if (lane == 0)
    for (int s = 0; s < p.S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const long long gw = (long long)blockIdx.x * nw + warp;
  const int *idx = p.idx + gw * (long long)p.chunks * p.R;
  const size_t row_floats = p.row_bytes / 4;
  unsigned long long issue_clk = 0;
  float4 acc = make_float4(0, 0, 0, 0);
  auto issue = [&](int c) {
    const int st = c % p.S;
    const uint32_t bar = smem_u32(bars + st), dst = smem_u32(ring + (size_t)st * stage_bytes);
    const int my_row = lane < p.R ? __ldg(idx + c * p.R + lane) : 0;
    if (p.lanes_issue == 1) {
      stage_idx[lane] = my_row;
      __syncwarp();
      if (elect_one()) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
        const long long t0 = clock64();
#pragma unroll 4
        for (int j = 0; j < p.R; ++j)
          bulk_row(dst + j * p.row_bytes, p.X + (size_t)stage_idx[j] * row_floats, p.row_bytes, bar);
        issue_clk += clock64() - t0;
      }
    } else {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(stage_bytes) : "memory");
      __syncwarp();
      const long long t0 = clock64();
      if (lane < p.R) bulk_row(dst + lane * p.row_bytes, p.X + (size_t)my_row * row_floats, p.row_bytes, bar);
      if (lane == 0) issue_clk += clock64() - t0;
    }
    __syncwarp();
  };
  for (int c = 0; c < p.S - 1 && c < p.chunks; ++c) issue(c);
  for (int c = 0; c < p.chunks; ++c) {
    if (c + p.S - 1 < p.chunks) issue(c + p.S - 1);
    const int st = c % p.S;
    while (!mbar_try(smem_u32(bars + st), (c / p.S) & 1)) {}
    const unsigned char *base = ring + (size_t)st * stage_bytes + lane * 16;
    for (int j = 0; j < p.R; ++j)
      for (int v = 0; v < p.row_bytes / 512; ++v) {
        const float4 x = *reinterpret_cast<const float4 *>(base + j * p.row_bytes + v * 512);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
    __syncwarp();
  }
  if (acc.x == 12345.f) p.out[gw] = acc.y + acc.z + acc.w;
  if (issue_clk) atomicAdd(p.clk, issue_clk);
}

int main() {
  const size_t bytes = 1536ull << 20;
  float *X, *out; int *idx; unsigned long long *clk;
  CK(cudaMalloc(&X, bytes)); CK(cudaMemset(X, 0, bytes));
  CK(cudaMalloc(&out, 1 << 22)); CK(cudaMalloc(&clk, 8));
  const size_t nidx = 16ull << 20;
  std::vector<int> h(nidx);
  CK(cudaMalloc(&idx, nidx * 4));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  for (int row_bytes : {512, 1024, 2048}) {
    const unsigned nrows = (unsigned)(bytes / row_bytes);
    unsigned s = 12345;
    for (size_t i = 0; i < nidx; ++i) { s = s * 1664525u + 1013904223u; h[i] = (int)((s >> 4) % nrows); }
    CK(cudaMemcpy(idx, h.data(), nidx * 4, cudaMemcpyHostToDevice));
    for (int lanes_issue : {1, 32})
      for (int ctas : {1, 2})
        for (int warps : {1, 2, 4, 8, 16})
          for (int R : {4, 8, 16, 32})
            for (int S : {2, 4, 8}) {
              if (lanes_issue == 32 && R > 32) continue;
              const size_t smem = 4096 + (size_t)warps * S * R * row_bytes;
              if (smem * ctas > 220 * 1024 || smem > 227 * 1024) continue;
              if ((size_t)R * row_bytes > 16384 && S > 2) continue;
              const int grid = prop.multiProcessorCount * ctas;
              const long long total_rows = 3ll << 20;     // rows per launch
              int chunks = (int)(total_rows / ((long long)grid * warps * R));
              if ((size_t)grid * warps * chunks * R > nidx) continue;
              P p{X, idx, out, clk, R, S, row_bytes, chunks, lanes_issue};
              cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
              float best = 1e9f;
              for (int it = 0; it < 3; ++it) {
                CK(cudaMemset(clk, 0, 8));
                cudaEventRecord(a);
                probe<<<grid, warps * 32, smem>>>(p);
                cudaEventRecord(b); CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, a, b);
                if (ms < best) best = ms;
              }
              unsigned long long hc = 0; CK(cudaMemcpy(&hc, clk, 8, cudaMemcpyDeviceToHost));
              const double rows = (double)grid * warps * chunks * R;
              printf("row %4d B  issue %-8s  %d CTA/SM x %2d warps  R=%2d S=%d  in flight/SM %6.1f KB  %8.1f us  %7.0f GB/s  %6.0f clk/UBLKCP\n",
                     row_bytes, lanes_issue == 1 ? "elected" : "per-lane", ctas, warps, R, S,
                     (double)ctas * warps * (S - 1) * R * row_bytes / 1024.0, best * 1e3, rows * row_bytes / best / 1e6,
                     (double)hc / rows);
              cudaEventDestroy(a); cudaEventDestroy(b);
              fflush(stdout);
            }
  }
  return 0;
}
