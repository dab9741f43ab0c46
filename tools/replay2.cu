// replay2.cu -- floor for the two gather stages: the workload's own address stream with a trivially lean
// kernel that also WRITES one output row per `per` gathered rows (stage A: ~4 gathers per Xe row; stage B:
// ~2 gathers per Y row), at several row widths and residencies.  No unit structure, no scales: only the
// traffic is real.  Usage: replay2 colind.bin  (int32 H^T column indices, tools/dump_colind.py)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int F, int PER, int STREAM>   // F floats per row (multiple of 128), PER gathers per stored row
__global__ void __launch_bounds__(256) lean(const int *__restrict__ ind, long long n, const float *__restrict__ src,
                                            float *__restrict__ dst) {
  constexpr int V = F / 128;
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  for (long long p = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * PER; p + PER <= n; p += nw * PER) {
    int v[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) v[u] = __ldg(ind + p + u);
    float4 x[PER][V];
#pragma unroll
    for (int u = 0; u < PER; ++u)
#pragma unroll
      for (int j = 0; j < V; ++j) x[u][j] = __ldg(reinterpret_cast<const float4 *>(src + (size_t)v[u] * F) + lane + 32 * j);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float4 s = x[0][j];
#pragma unroll
      for (int u = 1; u < PER; ++u) { s.x += x[u][j].x; s.y += x[u][j].y; s.z += x[u][j].z; s.w += x[u][j].w; }
      float *o = dst + (size_t)(p / PER) * F + (lane + 32 * j) * 4;
      if (STREAM) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
      else asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
    }
  }
}

// same traffic, but every warp owns BLOCKS of `BLK` consecutive positions claimed from a global counter (the
// work distribution of the library's stream kernel) instead of a fine round-robin interleave of all warps
template <int F, int PER, int BLK, int UNR>
__global__ void __launch_bounds__(256) lean_blocks(const int *__restrict__ ind, long long n, const float *__restrict__ src,
                                                   float *__restrict__ dst, int *counter) {
  constexpr int V = F / 128;
  const int lane = threadIdx.x & 31;
  for (;;) {
    int t = 0;
    if (lane == 0) t = atomicAdd(counter, 1);
    t = __shfl_sync(0xffffffffu, t, 0);
    const long long p0 = (long long)t * BLK;
    if (p0 >= n) break;
    float4 s[V];
#pragma unroll
    for (int j = 0; j < V; ++j) s[j] = make_float4(0, 0, 0, 0);
    for (long long p = p0; p < p0 + BLK && p + UNR <= n; p += UNR) {
      int v[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) v[u] = __ldg(ind + p + u);
      float4 x[UNR][V];
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int j = 0; j < V; ++j) x[u][j] = __ldg(reinterpret_cast<const float4 *>(src + (size_t)v[u] * F) + lane + 32 * j);
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
#pragma unroll
        for (int j = 0; j < V; ++j) { s[j].x += x[u][j].x; s[j].y += x[u][j].y; s[j].z += x[u][j].z; s[j].w += x[u][j].w; }
        if ((u + 1) % PER == 0) {
#pragma unroll
          for (int j = 0; j < V; ++j) {
            float *o = dst + (size_t)((p + u) / PER) * F + (lane + 32 * j) * 4;
            asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(s[j].x), "f"(s[j].y), "f"(s[j].z), "f"(s[j].w) : "memory");
            s[j] = make_float4(0, 0, 0, 0);
          }
        }
      }
    }
  }
}

template <int F, int PER, int BLK, int UNR>
void run_blocks(const char *name, const int *d_ind, long long n, const float *src, float *dst, int *counter) {
  for (int bps : {2, 3, 4, 6, 8}) {
    int g = 148 * bps;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float tot = 0;
    for (int it = 0; it < 6; ++it) {
      CK(cudaMemset(counter, 0, 4));
      cudaEventRecord(a);
      lean_blocks<F, PER, BLK, UNR><<<g, 256>>>(d_ind, n, src, dst, counter);
      cudaEventRecord(b); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (it) tot += ms;
    }
    printf("%-8s F=%3d per=%d blocks of %d, %d rows in flight, %d CTAs/SM  %8.1f us\n", name, F, PER, BLK, UNR, bps, tot / 5 * 1e3);
  }
}

template <int F, int PER, int STREAM>
void run(const char *name, const int *d_ind, long long n, const float *src, float *dst, size_t src_rows) {
  for (int bps : {2, 3, 4, 6, 8}) {
    int g = 148 * bps;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    lean<F, PER, STREAM><<<g, 256>>>(d_ind, n, src, dst); CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int it = 0; it < 5; ++it) lean<F, PER, STREAM><<<g, 256>>>(d_ind, n, src, dst);
    cudaEventRecord(b); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    double gathered = (double)n * F * 4, stored = (double)(n / PER) * F * 4, uniq = (double)src_rows * F * 4;
    printf("%-8s F=%3d per=%d cs=%d %d CTAs/SM  %8.1f us   rows moved %.2f GB (%.0f GB/s)   dram>= %.2f GB (%.0f GB/s)\n", name, F, PER,
           STREAM, bps, ms * 1e3, (gathered + stored) / 1e9, (gathered + stored) / ms / 1e6, (uniq + stored) / 1e9,
           (uniq + stored) / ms / 1e6);
  }
}

int main(int argc, char **argv) {
  const char *path = argc > 1 ? argv[1] : "gpurun_out/colind.bin";
  FILE *f = fopen(path, "rb");
  if (!f) { printf("cannot open %s\n", path); return 1; }
  fseek(f, 0, SEEK_END); long long nnz = ftell(f) / 4; fseek(f, 0, SEEK_SET);
  std::vector<int> h(nnz);
  if (fread(h.data(), 4, nnz, f) != (size_t)nnz) return 1;
  fclose(f);
  int N = 0; for (long long i = 0; i < nnz; ++i) N = h[i] + 1 > N ? h[i] + 1 : N;
  // stage-B stream: hyperedge ids, random inside the replica of the position (64 replicas of 7963 hyperedges)
  const int reps = 64, Mper = 7963, M = reps * Mper;
  std::vector<int> hb(nnz);
  unsigned s = 12345;
  for (long long i = 0; i < nnz; ++i) { s = s * 1664525u + 1013904223u; hb[i] = (int)(i * reps / nnz) * Mper + (int)((s >> 8) % Mper); }
  int *d, *db; float *X, *Y;
  CK(cudaMalloc(&d, nnz * 4)); CK(cudaMemcpy(d, h.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&db, nnz * 4)); CK(cudaMemcpy(db, hb.data(), nnz * 4, cudaMemcpyHostToDevice));
  const size_t maxF = 512;
  CK(cudaMalloc(&X, (size_t)N * maxF * 4)); CK(cudaMalloc(&Y, (size_t)N * maxF * 4));
  CK(cudaMemset(X, 0, (size_t)N * maxF * 4)); CK(cudaMemset(Y, 0, (size_t)N * maxF * 4));
  printf("nnz=%lld N=%d M=%d\n", nnz, N, M);
  // stage A: gather X rows (vertex ids), 4 per stored Xe row; stage B: gather Xe rows (hyperedge ids), 2 per stored Y row
  int *counter; CK(cudaMalloc(&counter, 4));
  run_blocks<128, 4, 64, 4>("A-blk", d, nnz, X, Y, counter);
  run_blocks<128, 4, 64, 8>("A-blk", d, nnz, X, Y, counter);
  run_blocks<128, 4, 16, 8>("A-blk", d, nnz, X, Y, counter);
  run_blocks<128, 2, 64, 8>("B-blk", db, nnz, X, Y, counter);
  run<128, 4, 0>("A", d, nnz, X, Y, N);
  run<128, 2, 1>("B", db, nnz, X, Y, M);
  run<128, 2, 0>("B", db, nnz, X, Y, M);
  run<512, 4, 0>("A", d, nnz, X, Y, N);
  run<512, 2, 1>("B", db, nnz, X, Y, M);
  return 0;
}
