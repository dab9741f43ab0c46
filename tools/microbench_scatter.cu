// microbench_scatter.cu -- which scatter primitive should stage 2 use on B200?
// Measures random-row (512 B rows, F=128 fp32) traffic for: plain v4 loads (gather), plain v4 stores,
// scalar atomicAdd, red.global.add.v4.f32, and TMA bulk reduce-add (cp.reduce.async.bulk) from smem,
// on an L2-resident (16 MB) and a DRAM-resident (2 GB) working set.   Build: see tools/Makefile.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int F = 128;
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

enum { GATHER, STORE, ATOMIC, REDV4, BULKRED, BULKLD, MIXED };

template <int MODE>
__global__ void __launch_bounds__(256) k(float *buf, uint32_t nrows, uint32_t rows_per_warp, float *sink, uint32_t nrows_y = 0, uint32_t ywin = 1) {
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * 8 + warp;
  float4 acc = make_float4(0, 0, 0, 0);
  float *my = smem + warp * 4 * F;                 // 4 row slots per warp
  __shared__ __align__(8) unsigned long long bar[8];
  if (MODE == BULKRED) {
    for (int s = 0; s < 4; ++s) reinterpret_cast<float4 *>(my + s * F)[lane] = make_float4(1, 1, 1, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
  }
  if (MODE == BULKLD) {
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[warp])));
    __syncwarp();
  }
  uint32_t phase = 0;
  for (uint32_t i = 0; i < rows_per_warp; i += 4) {
    uint32_t r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = mix(gw * rows_per_warp + i + u) % nrows;
    if (MODE == GATHER) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 x = __ldg(reinterpret_cast<const float4 *>(buf + (size_t)r[u] * F) + lane);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
    } else if (MODE == STORE) {
#pragma unroll
      for (int u = 0; u < 4; ++u) reinterpret_cast<float4 *>(buf + (size_t)r[u] * F)[lane] = make_float4(1, 1, 1, 1);
    } else if (MODE == ATOMIC) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 4; ++c) atomicAdd(buf + (size_t)r[u] * F + c * 32 + lane, 1.0f);
    } else if (MODE == REDV4) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(buf + (size_t)r[u] * F + lane * 4), "f"(1.0f) : "memory");
    } else if (MODE == MIXED) {   // the aggregation's mix: gather 4 rows (whole buffer), reduce, red to 4 rows (first 1/4)
      float4 sum = make_float4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 x = __ldg(reinterpret_cast<const float4 *>(buf + (size_t)r[u] * F) + lane);
        sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t y = nrows + (mix(r[u] ^ 0x9e3779b9u) % ywin);  // one window shared by all warps
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(buf + (size_t)y * F + lane * 4), "f"(sum.x), "f"(sum.y), "f"(sum.z), "f"(sum.w) : "memory");
      }
    } else if (MODE == BULKRED) {
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(buf + (size_t)r[u] * F),
                       "r"((uint32_t)__cvta_generic_to_shared(my + u * F)), "r"(F * 4) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      __syncwarp();
    } else if (MODE == BULKLD) {
      uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar[warp]);
      if (lane == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(4 * F * 4) : "memory");
#pragma unroll
        for (int u = 0; u < 4; ++u)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           (uint32_t)__cvta_generic_to_shared(my + u * F)), "l"(buf + (size_t)r[u] * F), "r"(F * 4), "r"(b) : "memory");
      }
      uint32_t done = 0;
      while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(b), "r"(phase) : "memory");
      phase ^= 1;
#pragma unroll
      for (int u = 0; u < 4; ++u) { float4 x = reinterpret_cast<float4 *>(my + u * F)[lane]; acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
      __syncwarp();
    }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) sink[0] = acc.x;
}

template <int MODE>
void run(const char *name, float *buf, uint32_t nrows, float *sink, const char *ws) {
  const uint32_t rows_per_warp = 512, blocks = 148 * 8 * 4;
  const size_t smem = 8 * 4 * F * sizeof(float);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<blocks, 256, smem>>>(buf, nrows, rows_per_warp, sink);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  for (int it = 0; it < 3; ++it) k<MODE><<<blocks, 256, smem>>>(buf, nrows, rows_per_warp, sink);
  cudaEventRecord(b); CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
  double bytes = (double)blocks * 8 * rows_per_warp * F * 4;
  printf("%-8s %-6s rows=%9u  %8.3f ms  %8.1f GB/s (row bytes moved)\n", name, ws, nrows, ms, bytes / ms / 1e6);
}

int main() {
  float *buf, *sink;
  const size_t big = (size_t)2 << 30;
  CK(cudaMalloc(&buf, big)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 0, big));
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t nrows = pass == 0 ? (16u << 20) / (F * 4) : (uint32_t)(big / (F * 4));
    const char *ws = pass == 0 ? "L2" : "DRAM";
    run<GATHER>("gather", buf, nrows, sink, ws);
    run<BULKLD>("bulk_ld", buf, nrows, sink, ws);
    run<STORE>("store", buf, nrows, sink, ws);
    run<ATOMIC>("atomic", buf, nrows, sink, ws);
    run<REDV4>("red.v4", buf, nrows, sink, ws);
    run<BULKRED>("bulk_red", buf, nrows, sink, ws);
  }
  // mixed: X = first 1 GB (DRAM-resident gathers); Y = second 1 GB, the reds of a warp land in a
  // sliding window of `ywin` rows (L2-resident window, as in the fused kernel where Y rows are
  // zero-filled just ahead of use)
  {
    const uint32_t nx = (uint32_t)((big / 2) / (F * 4)), ny = nx;
    CK(cudaFuncSetAttribute(k<MIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int occ : {8, 4, 2, 1})   // resident CTAs per SM, forced with dynamic shared memory
    for (uint32_t ywin : {40000u, 160000u, ny - 1}) {
      const uint32_t rows_per_warp = 512, blocks = 148 * 8 * 4;
      const size_t pad = occ >= 8 ? 0 : (size_t)(220 * 1024 / occ - 2048);
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      k<MIXED><<<blocks, 256, pad>>>(buf, nx, rows_per_warp, sink, ny, ywin);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(a);
      for (int it = 0; it < 3; ++it) k<MIXED><<<blocks, 256, pad>>>(buf, nx, rows_per_warp, sink, ny, ywin);
      cudaEventRecord(b); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= 3;
      double bytes = (double)blocks * 8 * rows_per_warp * F * 4;
      printf("mixed occ=%d CTAs/SM ywin=%8u rows  %8.3f ms  gather %8.1f GB/s + red %8.1f GB/s\n", occ, ywin, ms, bytes / ms / 1e6, bytes / ms / 1e6);
    }
  }
  return 0;
}
