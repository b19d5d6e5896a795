/*
 * ref_cpu_harness.cpp -- builds oracle/_ref/libcvr_ref_cpu.so: the REFERENCE's own
 * path kernels -- NaiveVolPTsk_kernel::d_render (NaiveVolPTsk_kernel.cuh:17-87) and
 * RegenerationVolPTsk_kernel::d_render_single_thread_regeneration
 * (RegenerationVolPTsk_kernel.cuh:146-232), with everything they call
 * (woodcockTracking, DeviceVolume::operator(), indexToCameraRay, AABB::intersect,
 * GGX::sample, HG::sample, atomicVectorAdd ...) -- compiled FOR THE HOST by plain g++
 * from the headers where they lie under /root/reference.  "The reference's
 * host-compiled __host__ __device__ estimator" of the north star: the CPU baseline of
 * bench.py (cpu_baseline.kind = "reference") and the whole-path pin of
 * oracle/cvr_oracle.c (tests/test_oracle.py compares per-path radiances bit for bit).
 *
 * TEST INFRASTRUCTURE ONLY; never on the product path.  This TU holds NO reference
 * code.  g++ sees CUDA's qualifiers as empty macros (host_defines.h without __CUDACC__),
 * so every __device__ function of the reference is an ordinary C++ function and a
 * __global__ kernel is a function called once per (virtual) CUDA thread.  What the CUDA
 * platform provides underneath those sources is supplied here, restated from the
 * platform's documented behaviour, not from the reference:
 *   - threadIdx / blockIdx / blockDim: thread_local variables set per call;
 *   - atomicAdd(float*) / atomicAdd(uint*): GCC __atomic builtins;
 *   - max(float, float): CUDA's overload (= fmaxf), declared BEFORE the reference is
 *     parsed -- helper_math.h's host section only has max(int, int), which would turn
 *     woodcockStep's max(u, EPSILON) (Utilities.cuh:134-136) into an integer maximum;
 *   - class Rng: the reference's (Rng.h:14-57) wraps cuRAND's device XORWOW and returns
 *     0 on the host; RNG_H_ is pre-defined and a class with the same interface over a
 *     host XORWOW (curand_kernel.h: _curand_init_scratch with subsequence 0 / offset 0,
 *     curand(), curand_uniform()) is supplied;
 *   - DeviceVolume<T>::get(x, y, z): the reference defines it as a point-sampled,
 *     clamp-addressed, unnormalised tex3D fetch (RenderKernelLauncher.cu:20-25,
 *     CudaVolPath.cpp:168-181); here the `volume_tex` handle carries a host pointer and
 *     the texture unit's clamp is min(index, size - 1) on the unsigned coordinate;
 *   - the six __constant__ symbols of RenderKernelLauncher.cu:67-72;
 *   - shim headers (oracle/_ref/shim_cpu, written by the Makefile): empty <cub/cub.cuh>
 *     (two CUB_PTX_* macros), <cooperative_groups.h> and "helper_cuda.h" stand-ins --
 *     only uninstantiated templates name them; empty glm headers as for the other harnesses.
 * One accommodation: Utilities.cuh:125 spells `__forceinline__ inline` (duplicate
 * specifier); __forceinline__ is defined empty for the reference includes.
 */
#include <cuda_runtime.h>

// libstdc++'s <math.h> / <stdlib.h> wrappers put the float overloads of sin, cos, tan, acos,
// atan2, sqrt, abs ... into the global namespace, which is what CUDA does for device code:
// GGX.h calls them unqualified with float arguments (GGX.h:95-162,244,284); with only the C
// declarations abs(float) would be the INTEGER abs and sin(float) a double evaluation.
#include <math.h>
#include <stdlib.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

// ---- CUDA platform pieces for host compilation (see header comment) -----------------
static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;

inline float atomicAdd(float* addr, float v) {
  uint32_t* p = reinterpret_cast<uint32_t*>(addr);
  uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED), want;
  float f;
  do {
    memcpy(&f, &old, 4);
    float s = f + v;
    memcpy(&want, &s, 4);
  } while (!__atomic_compare_exchange_n(p, &old, want, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return f;
}
inline unsigned int atomicAdd(unsigned int* addr, unsigned int v) {
  return __atomic_fetch_add(addr, v, __ATOMIC_RELAXED);
}
inline int atomicAdd(int* addr, int v) { return __atomic_fetch_add(addr, v, __ATOMIC_RELAXED); }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }  // CUDA's umax / umin overloads
inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
inline double max(double a, double b) { return fmax(a, b); }
inline double min(double a, double b) { return fmin(a, b); }
// CUDA math / intrinsics named only by reference code that is never called on this path
// (the unused GGX sampler GGX.h:183-209, the gradient medium Medium.h:30-106, atomicAggInc)
inline float norm3df(float a, float b, float c) { return sqrtf(a * a + b * b + c * c); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __popc(unsigned int x) { return __builtin_popcount(x); }
// warp / block primitives named by the reference's warp- and block-regeneration kernels,
// which are never instantiated here (REGENERATION_SYNCHRONIZATION_LEVEL 0, Defines.h:40-42)
inline int any(int p) { return p; }
inline int __any_sync(unsigned, int p) { return p; }
inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <class T>
inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T>
inline T __shfl(T v, int, int = 32) { return v; }
inline void __syncthreads() {}
inline int __syncthreads_or(int p) { return p; }
inline void __syncwarp(unsigned = 0xffffffffu) {}
inline void __threadfence_block() {}

// helper_math.h:51-68 has a "host implementations of CUDA functions" section for plain C++
// compilers (comparison-based fminf/fmaxf, single-precision rsqrtf).  The other reference
// checkers (oracle/_ref/libcvr_ref_host.so, tests/golden/) are host-compiled through nvcc,
// where that section is skipped and the CUDA toolkit's host definitions apply; the same
// definitions are used here so that all checkers agree bit for bit: libm's fminf / fmaxf,
// rsqrtf = (float)(1.0 / sqrt((double)x)) (crt/math_functions.hpp), max / min(int, int).
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline float rsqrtf(float a) { return (float)(1.0 / sqrt((double)a)); }
#define __CUDACC__
#include <helper_math.h>
#undef __CUDACC__

static bool g_dbg = false;  // diagnostics of refcpu_trace_one_naive (single-threaded)
static unsigned long long g_dbg_draws, g_dbg_gets;
#define RNG_H_
class Rng {  // interface of Rng.h:14-57 over cuRAND's XORWOW evaluated on the host
 public:
  struct State {
    unsigned int d, v[5];
  };
  Rng(State s) : s_(s) {}
  State getState() { return s_; }
  Rng(int seed = 1234) {
    // curand_init(seed, 0, 0, &state): the int seed converts to unsigned long long (sign-extended)
    const unsigned long long s = (unsigned long long)(long long)seed;
    const unsigned int s0 = (unsigned int)s ^ 0xaad26b49u, s1 = (unsigned int)(s >> 32) ^ 0xf7dcefddu;
    const unsigned int t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    s_.d = 6615241u + t1 + t0;
    s_.v[0] = 123456789u + t0;
    s_.v[1] = 362436069u ^ t0;
    s_.v[2] = 521288629u + t1;
    s_.v[3] = 88675123u ^ t1;
    s_.v[4] = 5783321u + t0;
  }
  unsigned int next() {
    if (g_dbg) ++g_dbg_draws;
    const unsigned int t = s_.v[0] ^ (s_.v[0] >> 2);
    s_.v[0] = s_.v[1], s_.v[1] = s_.v[2], s_.v[2] = s_.v[3], s_.v[3] = s_.v[4];
    s_.v[4] = (s_.v[4] ^ (s_.v[4] << 4)) ^ (t ^ (t << 1));
    s_.d += 362437u;
    return s_.v[4] + s_.d;
  }
  float getFloat() { return (float)next() * 2.3283064e-10f + (2.3283064e-10f / 2.0f); }  // curand_uniform: (0,1]
  uint getUint() { return (uint)getFloat(); }
  float2 getFloat2() {
    float a = getFloat();
    float b = getFloat();
    return make_float2(a, b);
  }
  float3 getFloat3() {
    float a = getFloat();
    float b = getFloat();
    float c = getFloat();
    return make_float3(a, b, c);
  }

 private:
  State s_;
};

#undef __forceinline__
#define __forceinline__

#include "Bsdf.h"
#include "CVRMath.h"
#include "Geometry.h"
#include "Medium.h"
#include "Ray.h"

// RenderKernelLauncher.cu:20-25 on the host: point fetch, clamp addressing
template <typename VolumeType>
VolumeType DeviceVolume<VolumeType>::get(uint x, uint y, uint z) {
  const uint cx = x < grid_resolution.x ? x : grid_resolution.x - 1;
  const uint cy = y < grid_resolution.y ? y : grid_resolution.y - 1;
  const uint cz = z < grid_resolution.z ? z : grid_resolution.z - 1;
  const VolumeType* h = reinterpret_cast<const VolumeType*>((uintptr_t)volume_tex);
  if (g_dbg) ++g_dbg_gets;
  return h[cx + (size_t)grid_resolution.x * (cy + (size_t)grid_resolution.y * cz)];
}

// RenderKernelLauncher.cu:67-72
float3x4 c_inv_view_mat;
float2 c_raster_to_view;
float2 c_resolution;
uint2 c_offset;
float2 c_pixel_index_range;
uint c_n_paths;

#include "NaiveVolPTsk_kernel.cuh"
#include "RegenerationVolPTsk_kernel.cuh"

typedef SimpleVolumeDeviceScene<DeviceMedium, GGX> RefScene;

static RefScene g_scene;
static uint g_tile_w, g_tile_h;

static void set_thread(uint tid) {
  blockDim = dim3(1, 1, 1), gridDim = dim3(1, 1, 1);
  threadIdx = make_uint3(0, 0, 0);
  blockIdx = make_uint3(tid, 0, 0);
}

template <class F>
static void run_threads(int n_threads, F f) {
  if (n_threads <= 1) {
    f(0);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(f, t);
  for (auto& t : th) t.join();
}

extern "C" {

/* density: x-fastest floats; albedo: x-fastest float4.  The arrays are BORROWED: they
 * must stay alive until the last refcpu_* render call. */
void refcpu_set_scene(const float* density, int dnx, int dny, int dnz, const float* albedo, int anx, int any_,
                      int anz, const float box_min[3], const float box_max[3], float scale, float max_density) {
  auto& m = g_scene.medium;
  m.density_volume.volume_tex = (cudaTextureObject_t)(uintptr_t)density;
  m.density_volume.grid_resolution = make_uint3(dnx, dny, dnz);
  m.albedo_volume.volume_tex = (cudaTextureObject_t)(uintptr_t)albedo;
  m.albedo_volume.grid_resolution = make_uint3(anx, any_, anz);
  m.max_density = max_density;
  m.scale = scale;
  m.density_AABB = AABB(make_float3(box_min[0], box_min[1], box_min[2]), make_float3(box_max[0], box_max[1], box_max[2]));
}

/* the constants CudaVolPath uploads (CudaVolPath.cpp:39-58,66-85,260-263) */
void refcpu_set_camera(const float inv_view[12], const float raster_to_view[2], unsigned tile_w, unsigned tile_h,
                       float full_w, float full_h, unsigned off_x, unsigned off_y) {
  for (int r = 0; r < 3; ++r)
    c_inv_view_mat.m[r] = make_float4(inv_view[4 * r], inv_view[4 * r + 1], inv_view[4 * r + 2], inv_view[4 * r + 3]);
  c_raster_to_view = make_float2(raster_to_view[0], raster_to_view[1]);
  c_resolution = make_float2((float)tile_w, (float)tile_h);
  c_pixel_index_range = make_float2(full_w, full_h);
  c_offset = make_uint2(off_x, off_y);
  g_tile_w = tile_w, g_tile_h = tile_h;
}

void refcpu_ggx_defaults(float alpha[2], float* eta) {
  alpha[0] = g_scene.bsdf.roughness.x, alpha[1] = g_scene.bsdf.roughness.y;
  *eta = g_scene.bsdf.int_ior_over_ext_ior;
}

/* naiveSK: one call of the reference's d_render per path id in [first, first + count);
 * `out` (tile_w * tile_h float4) is accumulated into through the reference's own
 * atomicVectorAdd. */
void refcpu_render_naive(unsigned long long first, unsigned long long count, float* out, int n_threads) {
  c_n_paths = 0xffffffffu;
  std::atomic<unsigned long long> head(0);
  run_threads(n_threads, [&](int) {
    for (;;) {
      const unsigned long long b = head.fetch_add(4096);
      if (b >= count) break;
      const unsigned long long e = b + 4096 < count ? b + 4096 : count;
      for (unsigned long long i = b; i < e; ++i) {
        set_thread((uint)(first + i));
        NaiveVolPTsk_kernel::d_render<RefScene>((float4*)out, g_scene);
      }
    }
  });
}

/* per-path radiance of the same: out_per_path[4 * i] = what path first + i added to its
 * pixel (0,0,0,0 when Russian roulette ended it, w = 1 when it escaped) */
void refcpu_trace_paths_naive(unsigned long long first, unsigned long long count, float* out_per_path, int n_threads) {
  c_n_paths = 0xffffffffu;
  const size_t npix = (size_t)g_tile_w * g_tile_h;
  std::atomic<unsigned long long> head(0);
  run_threads(n_threads, [&](int) {
    std::vector<float4> scratch(npix, make_float4(0, 0, 0, 0));
    for (;;) {
      const unsigned long long b = head.fetch_add(1024);
      if (b >= count) break;
      const unsigned long long e = b + 1024 < count ? b + 1024 : count;
      for (unsigned long long i = b; i < e; ++i) {
        const uint tid = (uint)(first + i);
        set_thread(tid);
        NaiveVolPTsk_kernel::d_render<RefScene>(scratch.data(), g_scene);
        float4& px = scratch[tid % npix];
        memcpy(out_per_path + 4 * i, &px, 16);
        px = make_float4(0, 0, 0, 0);
      }
    }
  });
}

/* regenerationSK (single-thread regeneration): n_threads persistent CUDA threads, one per
 * host thread, tid = 0 .. n_threads-1, sharing the reference's own `paths_head_global`
 * queue counter and `seed` (RegenerationVolPTsk_kernel.cuh:18-19).  With n_threads = 1 the
 * result is deterministic: stream Rng(seed) renders every path in path order. */
void refcpu_render_regen(unsigned long long n_paths, unsigned seed, float* out, int n_threads) {
  c_n_paths = (uint)n_paths;
  RegenerationVolPTsk_kernel::seed = seed;
  RegenerationVolPTsk_kernel::paths_head_global = 0;
  run_threads(n_threads < 1 ? 1 : n_threads, [&](int t) {
    set_thread((uint)t);
    RegenerationVolPTsk_kernel::d_render_single_thread_regeneration<RefScene>((float4*)out, g_scene);
  });
}

/* one naiveSK path with the number of uniforms drawn and texels fetched (test diagnostics) */
void refcpu_trace_one_naive(unsigned tid, float rgba[4], unsigned long long* n_draws, unsigned long long* n_texels) {
  c_n_paths = 0xffffffffu;
  const size_t npix = (size_t)g_tile_w * g_tile_h;
  std::vector<float4> scratch(npix, make_float4(0, 0, 0, 0));
  g_dbg_draws = g_dbg_gets = 0;
  g_dbg = true;
  set_thread(tid);
  NaiveVolPTsk_kernel::d_render<RefScene>(scratch.data(), g_scene);
  g_dbg = false;
  memcpy(rgba, &scratch[tid % npix], 16);
  *n_draws = g_dbg_draws, *n_texels = g_dbg_gets;
}

float refcpu_ggx_g1(const float alpha[2], const float v[3], const float m[3]) {
  return GGX_G1(make_float2(alpha[0], alpha[1]), make_float3(v[0], v[1], v[2]), make_float3(m[0], m[1], m[2]));
}
int refcpu_ggx_sample(const float wi[3], int seed, float wo[3], float* weight) {
  Rng rng(seed);
  float3 out = make_float3(wo[0], wo[1], wo[2]);
  bool ok = g_scene.bsdf.sample(make_float3(wi[0], wi[1], wi[2]), out, *weight, rng);
  wo[0] = out.x, wo[1] = out.y, wo[2] = out.z;
  return ok ? 1 : 0;
}

/* the reference's own device-only pieces, for unit pins */
float refcpu_density(const float p01[3]) {
  return g_scene.medium.density_volume(make_float3(p01[0], p01[1], p01[2]));
}
void refcpu_albedo_at_world(const float p[3], float rgba[4]) {
  float3 o = make_float3(p[0], p[1], p[2]);
  float4 a = g_scene.medium.sampleAlbedo(o);
  rgba[0] = a.x, rgba[1] = a.y, rgba[2] = a.z, rgba[3] = a.w;
}
void refcpu_camera_ray(unsigned image_id, int seed, float o[3], float d[3]) {
  Rng rng(seed);
  float2 pixel_index;
  pixel_index.x = (float)(image_id % ((uint)c_resolution.x)) + c_offset.x;
  pixel_index.y = (floorf((float)image_id / c_resolution.x)) + c_offset.y;
  Ray r = indexToCameraRay(pixel_index, c_pixel_index_range, c_raster_to_view, c_inv_view_mat, rng);
  o[0] = r.o.x, o[1] = r.o.y, o[2] = r.o.z, d[0] = r.d.x, d[1] = r.d.y, d[2] = r.d.z;
}
/* woodcockTracking (Utilities.cuh:138-155) through HeterogeneousMedium::sampleDistance
 * (Medium.h:135-143); returns the sampled distance, *n_draws = uniforms consumed */
float refcpu_woodcock(const float o[3], const float d[3], float max_t, int seed, int* scattered) {
  Rng rng(seed);
  float3 ro = make_float3(o[0], o[1], o[2]), rd = make_float3(d[0], d[1], d[2]);
  float t = 0.f;
  *scattered = g_scene.medium.sampleDistance(ro, rd, max_t, rng, t) ? 1 : 0;
  return t;
}

}  // extern "C"
