// Experiment (not product): how many DRAM sectors does one random 32-byte gather cost on B200, per load flavour?
// ncu on k_volpt_warp (fBm 1024^3) showed 3.4 DRAM sectors read and 3.9 L2 tag sectors per L1->L2 read request.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/exp/dram_granule tools/exp/dram_granule.cu
// run:   tools/exp/dram_granule [footprint_MiB] [l2_fetch_granularity|0]      (plain = timings; under ncu = sector counts)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

enum { V8_NC, V4_NC, F32_NC, V8_PLAIN, V4_CG, V8_NOALLOC, V8_L2_64, V8_L2_128, V8_L2_256, V4_CV, V8_EVICT_FIRST, V8_PAIR64, V8_QUAD128,
       BULK32, PREF_L2, PREF_L2_64, PREF_ONLY, N_FLAVOURS };
static const char* names[] = {"ld.global.nc.v8.f32", "ld.global.nc.v4.f32", "ld.global.nc.f32", "ld.global.v8.f32", "ld.global.cg.v4.f32",
                              "ld.global.nc.L1::no_allocate.v8", "ld.global.nc.L2::64B.v8", "ld.global.nc.L2::128B.v8", "ld.global.nc.L2::256B.v8",
                              "ld.volatile.global.v4.f32", "ld.global.nc.L2::evict_first.v8", "2 x v8 (aligned 64 B)", "4 x v8 (aligned 128 B)",
                              "cp.async.bulk 32 B -> smem", "prefetch.global.L2 64 ahead + ld.nc.v8", "same, loads carry L2::64B",
                              "prefetch.global.L2 only (fire and forget)"};

template <int FL>
__device__ __forceinline__ float load1(const float* p) {
  float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (FL == V8_NC) asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
  if (FL == V4_NC) asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  if (FL == F32_NC) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v[0]) : "l"(p));
  if (FL == V8_PLAIN) asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
  if (FL == V4_CG) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  if (FL == V8_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
  if (FL == V8_L2_64) asm volatile("ld.global.nc.L2::64B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
  if (FL == V8_L2_128) asm volatile("ld.global.nc.L2::128B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
  if (FL == V8_L2_256) asm volatile("ld.global.nc.L2::256B.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
  if (FL == V4_CV) asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "l"(p));
  if (FL == V8_EVICT_FIRST) asm volatile("ld.global.nc.L2::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
  return (v[0] + v[1]) + (v[2] + v[3]) + v[7];
}

__device__ __forceinline__ uint32_t hash(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// `per_thread` loads per thread, 8 in flight; every load at a hashed 32-byte cell of the buffer
template <int FL>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ cells, uint32_t n_cells, int per_thread, float* sink) {
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int i = 0; i < per_thread; i += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x = hash(x + 0x9e3779b9u);
      uint32_t cell = (uint32_t)(((unsigned long long)x * n_cells) >> 32);
      if (FL == V8_PAIR64) {
        cell &= ~1u;
        v[u] = load1<V8_NC>(cells + 8 * (size_t)cell) + load1<V8_NC>(cells + 8 * (size_t)cell + 8);
      } else if (FL == V8_QUAD128) {
        cell &= ~3u;
        v[u] = load1<V8_NC>(cells + 8 * (size_t)cell) + load1<V8_NC>(cells + 8 * (size_t)cell + 8) +
               load1<V8_NC>(cells + 8 * (size_t)cell + 16) + load1<V8_NC>(cells + 8 * (size_t)cell + 24);
      } else
        v[u] = load1<FL>(cells + 8 * (size_t)cell);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  if (acc == 123.456f) *sink = acc;
}

// the same gather through the bulk-copy engine: every thread copies its 32-byte cell into its own shared slot,
// 4 copies in flight per thread, one mbarrier per CTA phase
__global__ void __launch_bounds__(256) k_gather_bulk(const float* __restrict__ cells, uint32_t n_cells, int per_thread, float* sink) {
  __shared__ __align__(128) float stage[256 * 4 * 8];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(256));
  __syncthreads();
  float acc = 0.f;
  uint32_t phase = 0;
  for (int i = 0; i < per_thread; i += 4) {
    asm volatile("{ .reg .b64 t; mbarrier.arrive.expect_tx.shared::cta.b64 t, [%0], %1; }" ::"r"(bar_a), "r"(4 * 32) : "memory");
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      x = hash(x + 0x9e3779b9u);
      uint32_t cell = (uint32_t)(((unsigned long long)x * n_cells) >> 32);
      uint32_t dst = (uint32_t)__cvta_generic_to_shared(stage + (threadIdx.x * 4 + u) * 8);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(dst),
                   "l"(cells + 8 * (size_t)cell), "r"(bar_a)
                   : "memory");
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a), "r"(phase) : "memory");
    phase ^= 1;
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += stage[(threadIdx.x * 4 + u) * 8];
    __syncthreads();
  }
  if (acc == 123.456f) *sink = acc;
}

// software pipeline: the addresses of group i + DIST are prefetched into the L2 (no register, no scoreboard:
// the SM does not track a prefetch) while group i is loaded -- are DRAM-missing gathers bound by the SM's
// outstanding-request budget (then this is faster) or by the memory system itself (then it is not)?
template <int FL>
__global__ void __launch_bounds__(256) k_gather_pref(const float* __restrict__ cells, uint32_t n_cells, int per_thread, float* sink) {
  const uint32_t x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  constexpr int DIST = 8;  // groups of 8 loads ahead
  float acc = 0.f;
  uint32_t xp = x0, x = x0;
  for (int i = 0; i < DIST * 8 && FL != PREF_ONLY; ++i) {
    xp = hash(xp + 0x9e3779b9u);
    uint32_t cell = (uint32_t)(((unsigned long long)xp * n_cells) >> 32);
    asm volatile("prefetch.global.L2 [%0];" ::"l"(cells + 8 * (size_t)cell));
  }
  for (int i = 0; i < per_thread; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      xp = hash(xp + 0x9e3779b9u);
      uint32_t cell = (uint32_t)(((unsigned long long)xp * n_cells) >> 32);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(cells + 8 * (size_t)cell));
    }
    if (FL == PREF_ONLY) continue;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x = hash(x + 0x9e3779b9u);
      uint32_t cell = (uint32_t)(((unsigned long long)x * n_cells) >> 32);
      v[u] = FL == PREF_L2_64 ? load1<V8_L2_64>(cells + 8 * (size_t)cell) : load1<V8_NC>(cells + 8 * (size_t)cell);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  if (acc == 123.456f) *sink = acc;
}

template <int FL>
static void launch(const float* d, uint32_t n_cells, int per_thread, float* sink, int grid) {
  if (FL == BULK32) k_gather_bulk<<<grid, 256>>>(d, n_cells, per_thread, sink);
  else if (FL == PREF_L2 || FL == PREF_L2_64 || FL == PREF_ONLY) k_gather_pref<FL><<<grid, 256>>>(d, n_cells, per_thread, sink);
  else k_gather<FL><<<grid, 256>>>(d, n_cells, per_thread, sink);
}
typedef void (*launch_fn)(const float*, uint32_t, int, float*, int);
template <int... I> struct Seq {};
template <int N, int... I> struct Gen : Gen<N - 1, N - 1, I...> {};
template <int... I> struct Gen<0, I...> { typedef Seq<I...> type; };
template <int... I> static void fill(launch_fn* t, Seq<I...>) { launch_fn a[] = {launch<I>...}; for (int i = 0; i < (int)sizeof...(I); ++i) t[i] = a[i]; }

int main(int argc, char** argv) {
  size_t mib = argc > 1 ? atoll(argv[1]) : 8192;
  int gran = argc > 2 ? atoi(argv[2]) : 0;
  int only = argc > 3 ? atoi(argv[3]) : -1;
  if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran));
  size_t got = 0;
  CK(cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity));
  size_t bytes = mib << 20;
  float *d, *sink;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(d, 0, bytes));
  uint32_t n_cells = (uint32_t)(bytes / 32);
  launch_fn tab[N_FLAVOURS];
  fill(tab, Gen<N_FLAVOURS>::type());
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms * 16, per_thread = 128;
  const double loads = (double)grid * 256 * per_thread;
  printf("footprint %zu MiB, cudaLimitMaxL2FetchGranularity %zu, %d SMs, %.1f M gathers per launch\n", mib, got, sms, loads / 1e6);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int f = 0; f < N_FLAVOURS; ++f) {
    if (only >= 0 && f != only) continue;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      tab[f](d, n_cells, per_thread, sink, grid);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best) best = ms;
    }
    double mult = f == V8_PAIR64 ? 2 : f == V8_QUAD128 ? 4 : 1;
    printf("%-36s %8.3f ms  %7.2f G gathers/s  %8.1f GB/s of requested 32-B sectors\n", names[f], best, loads / best / 1e6,
           loads * 32 * mult / best / 1e6);
  }
  return 0;
}
