// Micro-benchmark behind the roofline discussion in DESIGN.md: what does a B200 deliver for the access pattern of the
// FM forward -- random 128-byte parameter rows gathered by id, ids streamed -- as a function of the table size (L2-resident
// .. far larger than L2), with and without L2 eviction hints?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31);
}
__global__ void fill_ids(uint32_t* ids, int64_t n, uint32_t rows) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) ids[i] = (uint32_t)(splitmix64(i) % rows); }
__global__ void fill_f(float* a, int64_t n) { int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = 1.0f; }

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, int64_t n)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// HINT 0: plain ld.global.nc; 1: table rows evict_last (fraction frac), ids evict_first
template <int U, int HINT>
__global__ void __launch_bounds__(256, 4) gather_kernel(const uint32_t* __restrict__ ids, int64_t n_ids, const float* __restrict__ table, float* __restrict__ out, float frac)
{
  const int lane = threadIdx.x & 31, g = lane >> 3, l = lane & 7;
  uint64_t pol_tab = 0, pol_ids = 0;
  if (HINT) {
    asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_unchanged.b64 %0, %1;" : "=l"(pol_tab) : "f"(frac));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_ids));
  }
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int64_t base = warp * (4 * U); base + 4 * U <= n_ids; base += nwarps * (4 * U)) {
    uint32_t c[U];
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t* p = ids + base + u * 4 + g;
      if (HINT) asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(c[u]) : "l"(p), "l"(pol_ids));
      else asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(c[u]) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4* p = reinterpret_cast<const float4*>(table + (size_t)c[u] * 32) + l;
      if (HINT) asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p), "l"(pol_tab));
      else asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == -1.f) out[0] = acc.x;
}

template <class F> float time_ms(F f, int reps)
{
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  f(); f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
  float ms; CK(cudaEventElapsedTime(&ms, a, b));
  CK(cudaGetLastError());
  return ms / reps;
}

int main()
{
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, L2 %d MB, persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB\n", prop.name, prop.multiProcessorCount,
         prop.l2CacheSize >> 20, prop.persistingL2CacheMaxSize >> 20, prop.accessPolicyMaxWindowSize >> 20);
  const int64_t n_ids = 390000000ll / 2;      // half of configs[1]'s gathers per launch
  uint32_t* ids; float* out; CK(cudaMalloc(&ids, n_ids * 4)); CK(cudaMalloc(&out, 64));
  {
    const int64_t nf = 1ll << 28;             // 1 GiB copy
    float *a, *b; CK(cudaMalloc(&a, nf * 4)); CK(cudaMalloc(&b, nf * 4));
    fill_f<<<(unsigned)((nf + 255) / 256), 256>>>(a, nf);
    float ms = time_ms([&] { copy_kernel<<<148 * 8, 256>>>((const float4*)a, (float4*)b, nf / 4); }, 10);
    printf("copy 1 GiB: %.3f ms  %.0f GB/s (read+write)\n", ms, 2.0 * nf * 4 / ms / 1e6);
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  const int sizes_mb[] = {16, 64, 96, 128, 192, 256, 1024, 6400};
  for (int mb : sizes_mb) {
    const uint32_t rows = (uint32_t)((size_t)mb * 1024 * 1024 / 128);
    float* table; CK(cudaMalloc(&table, (size_t)rows * 128));
    fill_f<<<(unsigned)(((int64_t)rows * 32 + 255) / 256), 256>>>(table, (int64_t)rows * 32);
    fill_ids<<<(unsigned)((n_ids + 255) / 256), 256>>>(ids, n_ids, rows);
    CK(cudaDeviceSynchronize());
    const int grid = prop.multiProcessorCount * 4;
    float t4 = time_ms([&] { gather_kernel<4, 0><<<grid, 256>>>(ids, n_ids, table, out, 0.f); }, 5);
    float t8 = time_ms([&] { gather_kernel<8, 0><<<grid, 256>>>(ids, n_ids, table, out, 0.f); }, 5);
    float h50 = time_ms([&] { gather_kernel<4, 1><<<grid, 256>>>(ids, n_ids, table, out, 0.5f); }, 5);
    float h75 = time_ms([&] { gather_kernel<4, 1><<<grid, 256>>>(ids, n_ids, table, out, 0.75f); }, 5);
    float h100 = time_ms([&] { gather_kernel<4, 1><<<grid, 256>>>(ids, n_ids, table, out, 1.0f); }, 5);
    const double gb = (double)n_ids * 132 / 1e9;     // 128-byte row + 4-byte id per gather, L2 -> SM
    printf("table %5d MB: U4 %.3f ms (%.0f GB/s L2->SM, %.2f Grows/s) | U8 %.3f ms | hint .5 %.3f  .75 %.3f  1.0 %.3f ms\n", mb, t4, gb / t4 * 1e3,
           n_ids / t4 / 1e6, t8, h50, h75, h100);
    CK(cudaFree(table));
  }
  return 0;
}
