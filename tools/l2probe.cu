// l2probe.cu -- how long does a written line survive in the B200 L2 under streaming traffic, and do the
// eviction-priority hints / discard change that?  (Design input for handing the hyperedge features from
// stage A to stage B through the L2 inside one launch.)
//
// One persistent launch, static round robin (warp w takes slices w, w + W, ...).  Slice i:
//   reads 5 X rows (512 B each), writes 2 Xe rows      -- "stage A" of slice i
//   reads the 2 Xe rows of slice i - L, writes 5 Y rows -- "stage B" of slice i - L
// so 7 KB pass through the L2 per slice (5 KB of it compulsory DRAM traffic) and an Xe line is read back after
// L x 7 KB of other traffic.  L >= 2 W, so the line was written a full round earlier.  Reported per
// configuration: time, and (under `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`) the DRAM bytes:
// Xe reads that missed = read - X bytes, Xe write-backs = write - Y bytes.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

enum { P_NONE = 0, P_FIRST = 1, P_LAST = 2, P_NORMAL = 3 };

__device__ __forceinline__ uint64_t make_policy(int kind) {
  uint64_t p = 0;
  if (kind == P_FIRST) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else if (kind == P_LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ld_hint(const float *p, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float4 ld_cg(const float *p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_hint(float *p, float4 v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_cs(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct Cfg {
  int px, pxe_w, pxe_r, py;   // eviction-priority policy of each access class (0 = plain instruction)
  int discard;                // discard.global.L2 the Xe lines after their read-back
  int lag;                    // slices between the write of an Xe row and its read-back
  int nslice;
};

constexpr int kRow = 128;     // floats per row (512 B)

__global__ void __launch_bounds__(256) probe(const float *__restrict__ X, float *Xe, float *__restrict__ Y, Cfg c) {
  const int lane = threadIdx.x & 31;
  const uint64_t polx = make_policy(c.px), polew = make_policy(c.pxe_w), poler = make_policy(c.pxe_r), poly = make_policy(c.py);
  const int W = gridDim.x * (blockDim.x >> 5);
  for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < c.nslice + c.lag; i += W) {
    float4 v[5], a, b;
    const bool doA = i < c.nslice, doB = i >= c.lag;
    if (doA) {
      const float *x = X + (long long)i * 5 * kRow + lane * 4;
#pragma unroll
      for (int u = 0; u < 5; ++u) v[u] = c.px ? ld_hint(x + u * kRow, polx) : __ldg(reinterpret_cast<const float4 *>(x + u * kRow));
    }
    float *eb = Xe + (long long)(i - c.lag) * 2 * kRow + lane * 4;
    if (doB) {
      if (c.pxe_r) { a = ld_hint(eb, poler); b = ld_hint(eb + kRow, poler); }
      else { a = ld_cg(eb); b = ld_cg(eb + kRow); }
    }
    if (doA) {
      float *e = Xe + (long long)i * 2 * kRow + lane * 4;
      const float4 s0 = make_float4(v[0].x + v[1].x + v[2].x, v[0].y + v[1].y + v[2].y, v[0].z + v[1].z + v[2].z, v[0].w + v[1].w + v[2].w);
      const float4 s1 = make_float4(v[3].x + v[4].x, v[3].y + v[4].y, v[3].z + v[4].z, v[3].w + v[4].w);
      if (c.pxe_w) { st_hint(e, s0, polew); st_hint(e + kRow, s1, polew); }
      else { *reinterpret_cast<float4 *>(e) = s0; *reinterpret_cast<float4 *>(e + kRow) = s1; }
    }
    if (doB) {
      float *y = Y + (long long)(i - c.lag) * 5 * kRow + lane * 4;
#pragma unroll
      for (int u = 0; u < 5; ++u) {
        const float4 r = u < 3 ? a : b;
        if (c.py) st_hint(y + u * kRow, r, poly); else st_cs(y + u * kRow, r);
      }
      if (c.discard) {
        __syncwarp();
        if ((lane & 7) == 0) {
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(eb) : "memory");
          asm volatile("discard.global.L2 [%0], 128;" ::"l"(eb + kRow) : "memory");
        }
      }
    }
  }
}

int main(int argc, char **argv) {
  const bool ncu = argc > 1 && !strcmp(argv[1], "ncu");
  const long long x_bytes = 1ll << 30;
  const int nslice = (int)(x_bytes / (5 * kRow * 4));
  float *X, *Xe, *Y;
  CK(cudaMalloc(&X, x_bytes)); CK(cudaMalloc(&Y, x_bytes)); CK(cudaMalloc(&Xe, x_bytes / 5 * 2 + (1 << 20)));
  CK(cudaMemset(X, 0, x_bytes)); CK(cudaMemset(Y, 0, x_bytes)); CK(cudaMemset(Xe, 0, x_bytes / 5 * 2));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("L2 %d MB, persisting max %d MB, SMs %d; X = Y = %.3f GB, Xe = %.3f GB\n", prop.l2CacheSize >> 20,
         prop.persistingL2CacheMaxSize >> 20, prop.multiProcessorCount, x_bytes / 1e9, x_bytes * 0.4 / 1e9);

  struct Pol { const char *name; int px, pw, pr, py, disc; };
  const Pol pols[] = {
      {"plain (ldg / st / ld.cg / st.cs)", 0, 0, 0, 0, 0},
      {"plain + discard", 0, 0, 0, 0, 1},
      {"X first, Xe-w last, Xe-r normal, Y first", P_FIRST, P_LAST, P_NORMAL, P_FIRST, 0},
      {"X first, Xe-w last, Xe-r first, Y first", P_FIRST, P_LAST, P_FIRST, P_FIRST, 0},
      {"X first, Xe-w last, Xe-r first, Y first + discard", P_FIRST, P_LAST, P_FIRST, P_FIRST, 1},
      {"X first, Xe-w normal, Xe-r normal, Y first + discard", P_FIRST, P_NORMAL, P_NORMAL, P_FIRST, 1},
      {"X first, Xe plain, Y first + discard", P_FIRST, 0, 0, P_FIRST, 1},
  };
  const int npol = sizeof(pols) / sizeof(pols[0]);
  for (int wpc : {8, 4}) {          // warps per CTA, two CTAs per SM
    const int ctas = prop.multiProcessorCount * 2;
    const int W = ctas * wpc;
    for (int mult : {2, 4, 8, 16, 32, 1 << 20}) {
      const int lag = mult == (1 << 20) ? nslice : mult * W;
      for (int pi = 0; pi < npol; ++pi) {
        const Pol &po = pols[pi];
        Cfg c{po.px, po.pw, po.pr, po.py, po.disc, lag, nslice};
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9f;
        const int reps = ncu ? 1 : 4;
        for (int it = 0; it < reps; ++it) {
          cudaEventRecord(a);
          probe<<<ctas, wpc * 32>>>(X, Xe, Y, c);
          cudaEventRecord(b); CK(cudaDeviceSynchronize());
          float ms; cudaEventElapsedTime(&ms, a, b);
          if (it || reps == 1) best = ms < best ? ms : best;
        }
        cudaEventDestroy(a); cudaEventDestroy(b);
        printf("warps %5d  lag %8d slices = %7.1f MB of L2 traffic  %-52s %8.1f us  alg %6.0f GB/s\n", W, lag, lag * 7168.0 / 1048576.0,
               po.name, best * 1e3, 2.0 * x_bytes / best / 1e6);
        fflush(stdout);
      }
    }
  }
  return 0;
}
