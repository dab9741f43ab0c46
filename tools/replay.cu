// replay.cu -- how fast can the ADDRESS STREAM of the real workload go with a trivially lean kernel?
// Reads the column-index list of a hypergraph (int32 file written by tools/dump_colind.py) and replays
// its memory accesses at F = 128 with the microbenchmark's loop shape: every warp takes 4 consecutive
// index positions, gathers those 4 X rows (LDG.128), sums them, and issues 4 red.v4 to the same 4 Y rows.
// No segment structure, no scales, no zero-fill (results are meaningless; only the traffic is real).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
constexpr int F = 128;

template <int MODE>  // 0 gather+red, 1 gather only, 2 red only
__global__ void __launch_bounds__(256) replay(const int *__restrict__ colind, long long nnz, const float *__restrict__ X,
                                             float *Y, float *sink) {
  const int lane = threadIdx.x & 31;
  const long long nw = (long long)gridDim.x * 8;
  float4 keep = make_float4(0, 0, 0, 0);
  for (long long p = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 4; p < nnz; p += nw * 4) {
    int v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = p + u < nnz ? __ldg(colind + p + u) : 0;
    float4 sum = make_float4(1, 1, 1, 1);
    if (MODE != 2) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 x = __ldg(reinterpret_cast<const float4 *>(X + (size_t)v[u] * F) + lane);
        sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
      }
    }
    if (MODE != 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(Y + (size_t)v[u] * F + lane * 4), "f"(sum.x), "f"(sum.y), "f"(sum.z), "f"(sum.w) : "memory");
    } else { keep.x += sum.x; }
  }
  if (keep.x == 12345.678f) sink[0] = keep.x;
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
  int *d; float *X, *Y, *sink;
  CK(cudaMalloc(&d, nnz * 4)); CK(cudaMemcpy(d, h.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&X, (size_t)N * F * 4)); CK(cudaMalloc(&Y, (size_t)N * F * 4)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(X, 0, (size_t)N * F * 4)); CK(cudaMemset(Y, 0, (size_t)N * F * 4));
  printf("nnz=%lld N=%d F=%d  rows gathered+reduced: %.2f GB each\n", nnz, N, F, nnz * 512.0 / 1e9);
  for (int blocks_per_sm : {8, 4, 2}) {
    for (int mode = 0; mode < 3; ++mode) {
      cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
      auto run = [&]() {
        int g = 148 * blocks_per_sm;
        if (mode == 0) replay<0><<<g, 256>>>(d, nnz, X, Y, sink);
        if (mode == 1) replay<1><<<g, 256>>>(d, nnz, X, Y, sink);
        if (mode == 2) replay<2><<<g, 256>>>(d, nnz, X, Y, sink);
      };
      run(); CK(cudaDeviceSynchronize());
      cudaEventRecord(a);
      for (int it = 0; it < 5; ++it) run();
      cudaEventRecord(b); CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
      printf("grid=%4d (%d CTAs/SM) %-11s %8.1f us  (%.0f GB/s of rows)\n", 148 * blocks_per_sm, blocks_per_sm,
             mode == 0 ? "gather+red" : mode == 1 ? "gather" : "red", ms * 1e3, nnz * 512.0 * (mode == 0 ? 2 : 1) / ms / 1e6);
    }
  }
  return 0;
}
