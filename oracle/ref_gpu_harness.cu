/*
 * ref_gpu_harness.cu -- builds oracle/_ref/libcvr_ref_gpu.so: the REFERENCE's own
 * naiveSK and regenerationSK(thread) kernels, instantiated from the reference's
 * headers where they lie under /root/reference, compiled for sm_100a.
 *
 * TEST INFRASTRUCTURE ONLY ("the reference on this box": parity checker and GPU
 * baseline; never on the product path).  This TU holds NO reference code, only:
 *  - the glue the reference keeps in RenderKernelLauncher.cu, re-declared because
 *    that file cannot be compiled here (it pulls Config.h -> Camera.h -> glm and
 *    the CUB kernels with MSVC-only syntax): the six __constant__ symbols the
 *    kernels name (RenderKernelLauncher.cu:67-72) and DeviceVolume<T>::get as a
 *    point-sampled tex3D fetch (RenderKernelLauncher.cu:20-25);
 *  - texture setup equal to CudaVolPath.cpp:147-181 (cudaArray, point filter,
 *    clamp, unnormalised coordinates, element read mode);
 *  - launch configuration equal to Occupancy.cuh:42-70 + RenderKernelLauncher.cu
 *    :147-158 (naive) and :280-290,331-335 (regeneration).
 * One accommodation: Utilities.cuh:125 spells `__forceinline__ inline`, which nvcc
 * rejects as a duplicate specifier; __forceinline__ is re-defined WITHOUT its own
 * `inline` for the duration of the reference includes (no copy of the tree).
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>

#include <helper_math.h>

#pragma push_macro("__forceinline__")
#undef __forceinline__
#ifdef __CUDA_ARCH__
#define __forceinline__ __attribute__((always_inline))
#else
#define __forceinline__
#endif

#include "Bsdf.h"
#include "CVRMath.h"
#include "Geometry.h"
#include "Medium.h"
#include "Ray.h"
#include "Rng.h"

// RenderKernelLauncher.cu:20-25 equivalent (must precede the kernels' use)
template <typename VolumeType>
__device__ VolumeType DeviceVolume<VolumeType>::get(uint x, uint y, uint z) {
  return tex3D<VolumeType>(volume_tex, x, y, z);
}

// RenderKernelLauncher.cu:67-72
__constant__ float3x4 c_inv_view_mat;
__constant__ float2 c_raster_to_view;
__constant__ float2 c_resolution;
__constant__ uint2 c_offset;
__constant__ float2 c_pixel_index_range;
__constant__ uint c_n_paths;

#include "NaiveVolPTsk_kernel.cuh"
#include "RegenerationVolPTsk_kernel.cuh"
// The streamingSK kernel: a syntax-patched COPY of StreamingVolPTsk_kernel.cuh / MortonSort.h that
// oracle/patch_ref_streaming.py writes into oracle/_ref/patched/ at build time (typename and
// class-scope-specialisation fixes only); that directory precedes the reference's
// on the include path for these two files.
#ifdef CVR_REF_STREAMING
#include "StreamingVolPTsk_kernel.cuh"
#endif

#pragma pop_macro("__forceinline__")

typedef SimpleVolumeDeviceScene<DeviceMedium, GGX> RefScene;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      snprintf(g_err, sizeof g_err, "%s: %s", #x, cudaGetErrorString(e_));         \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

static char g_err[512];

struct RefState {
  cudaArray_t d_density = nullptr, d_albedo = nullptr;
  cudaTextureObject_t t_density = 0, t_albedo = 0;
  RefScene scene;
  float4* d_out = nullptr;
  size_t out_pixels = 0;
  uint seed = 0;
  uint n_paths = 0;
  uint tile_w = 0, tile_h = 0;
  Threads d_threads;  // streamingSK queue slice (RenderKernelLauncher.cu:522-538)
  size_t n_threads = 0;
};
static RefState g;

template <class T>
static int make_tex(const T* h, int nx, int ny, int nz, cudaArray_t* arr, cudaTextureObject_t* tex) {
  cudaChannelFormatDesc desc = cudaCreateChannelDesc<T>();
  cudaExtent ext = make_cudaExtent(nx, ny, nz);
  CK(cudaMalloc3DArray(arr, &desc, ext));
  cudaMemcpy3DParms p;
  memset(&p, 0, sizeof p);
  p.srcPtr = make_cudaPitchedPtr((void*)h, nx * sizeof(T), nx, ny);
  p.dstArray = *arr;
  p.extent = ext;
  p.kind = cudaMemcpyHostToDevice;
  CK(cudaMemcpy3D(&p));
  cudaResourceDesc res;
  memset(&res, 0, sizeof res);
  res.resType = cudaResourceTypeArray;
  res.res.array.array = *arr;
  cudaTextureDesc td;
  memset(&td, 0, sizeof td);
  td.normalizedCoords = false;
  td.filterMode = cudaFilterModePoint;
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.readMode = cudaReadModeElementType;
  CK(cudaCreateTextureObject(tex, &res, &td, nullptr));
  return 0;
}

extern "C" {

const char* refgpu_last_error() { return g_err; }

int refgpu_release() {
  if (g.t_density) cudaDestroyTextureObject(g.t_density);
  if (g.t_albedo) cudaDestroyTextureObject(g.t_albedo);
  if (g.d_density) cudaFreeArray(g.d_density);
  if (g.d_albedo) cudaFreeArray(g.d_albedo);
  if (g.d_out) cudaFree(g.d_out);
  if (g.n_threads) {
    cudaFree(g.d_threads.paths.rays.o), cudaFree(g.d_threads.paths.rays.d);
    cudaFree(g.d_threads.paths.throughputs), cudaFree(g.d_threads.image_ids);
  }
  g = RefState();
  return 0;
}

/* density: x-fastest floats; albedo: x-fastest float4 */
int refgpu_set_scene(const float* density, int dnx, int dny, int dnz, const float* albedo,
                     int anx, int any, int anz, const float box_min[3], const float box_max[3],
                     float scale, float max_density) {
  refgpu_release();
  if (make_tex<float>(density, dnx, dny, dnz, &g.d_density, &g.t_density)) return 1;
  if (make_tex<float4>((const float4*)albedo, anx, any, anz, &g.d_albedo, &g.t_albedo)) return 1;
  auto& m = g.scene.medium;
  m.density_volume.volume_tex = g.t_density;
  m.density_volume.grid_resolution = make_uint3(dnx, dny, dnz);
  m.albedo_volume.volume_tex = g.t_albedo;
  m.albedo_volume.grid_resolution = make_uint3(anx, any, anz);
  m.max_density = max_density;
  m.scale = scale;
  m.density_AABB = AABB(make_float3(box_min[0], box_min[1], box_min[2]),
                        make_float3(box_max[0], box_max[1], box_max[2]));
  return 0;
}

int refgpu_set_camera(const float inv_view[12], const float raster_to_view[2], unsigned tile_w,
                      unsigned tile_h, float full_w, float full_h, unsigned off_x, unsigned off_y) {
  float2 rtv = make_float2(raster_to_view[0], raster_to_view[1]);
  float2 res = make_float2((float)tile_w, (float)tile_h);
  float2 range = make_float2(full_w, full_h);
  uint2 off = make_uint2(off_x, off_y);
  CK(cudaMemcpyToSymbol(c_inv_view_mat, inv_view, sizeof(float4) * 3));
  CK(cudaMemcpyToSymbol(c_raster_to_view, &rtv, sizeof rtv));
  CK(cudaMemcpyToSymbol(c_resolution, &res, sizeof res));
  CK(cudaMemcpyToSymbol(c_pixel_index_range, &range, sizeof range));
  CK(cudaMemcpyToSymbol(c_offset, &off, sizeof off));
  if ((size_t)tile_w * tile_h != g.out_pixels) {
    if (g.d_out) cudaFree(g.d_out);
    g.out_pixels = (size_t)tile_w * tile_h;
    CK(cudaMalloc(&g.d_out, g.out_pixels * sizeof(float4)));
  }
  g.tile_w = tile_w, g.tile_h = tile_h;
  return 0;
}

/* kernel: 0 = NaiveVolPTsk_kernel::d_render, 1 = regeneration single-thread,
 * 2 = StreamingVolPTsk_kernel::d_render as the launcher instantiates it at HEAD (VARIANT
 * defaulted = kSortingRays, Q16), 3 = the same with VARIANT = kClassic (scan compaction).
 * Renders `iterations` spp of the current tile into a zeroed buffer, copies the raw
 * accumulation (NOT divided) to host_out (tile_w*tile_h float4), reports the kernel
 * time in ms (CUDA events) and the launch shape used.  `seed` is the regeneration
 * kernel's `seed` symbol (the launcher's seed_ before this launch). */
int refgpu_render(int kernel, unsigned iterations, unsigned seed, float* host_out, float* ms,
                  int* grid_out, int* block_out) {
  uint n_paths = g.tile_w * g.tile_h * iterations; /* uint, as RenderKernelLauncher.cu:125 */
  CK(cudaMemcpyToSymbol(c_n_paths, &n_paths, sizeof n_paths));
  CK(cudaMemset(g.d_out, 0, g.out_pixels * sizeof(float4)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  int min_grid = 0, block = 0;
  if (kernel == 0) {
    auto k = NaiveVolPTsk_kernel::d_render<RefScene>;
    CK(cudaOccupancyMaxPotentialBlockSize(&min_grid, &block, k, 0, 0));
    int grid = (int)((n_paths + (uint)block - 1) / (uint)block);
    CK(cudaEventRecord(e0));
    k<<<grid, block>>>(g.d_out, g.scene);
    CK(cudaEventRecord(e1));
    *grid_out = grid;
  } else if (kernel == 2 || kernel == 3) {
#ifdef CVR_REF_STREAMING
    // StreamingVolPTsk::init / allocateDeviceMemory / launchRender / reset (RenderKernelLauncher.cu:486-575):
    // block = STREAMING_THREADS_BLOCK, grid = occupancy x SMs (one block per SM when the shared memory
    // would not fit), queue of grid*block*ITEMS slots, d_paths_head_global = 0, c_seed = seed
    void* k = kernel == 2
                  ? (void*)StreamingVolPTsk_kernel::d_render<STREAMING_THREADS_BLOCK, STREAMING_ITEMS_PER_THREAD, RefScene>
                  : (void*)StreamingVolPTsk_kernel::d_render<STREAMING_THREADS_BLOCK, STREAMING_ITEMS_PER_THREAD, RefScene,
                                                             StreamingVolPTsk_kernel::kClassic>;
    block = STREAMING_THREADS_BLOCK;
    int per_sm = 0, dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDevice(&dev));
    CK(cudaGetDeviceProperties(&prop, dev));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, block, 0));
    int grid = per_sm * prop.multiProcessorCount;
    if (prop.sharedMemPerMultiprocessor < STREAMING_SHARED_MEMORY * (float)grid / (float)prop.multiProcessorCount)
      grid = prop.multiProcessorCount;
    size_t n_threads = (size_t)grid * block * STREAMING_ITEMS_PER_THREAD;
    if (n_threads != g.n_threads) {
      if (g.n_threads) {
        cudaFree(g.d_threads.paths.rays.o), cudaFree(g.d_threads.paths.rays.d);
        cudaFree(g.d_threads.paths.throughputs), cudaFree(g.d_threads.image_ids);
      }
      CK(cudaMalloc(&g.d_threads.paths.rays.o, n_threads * sizeof(float3)));
      CK(cudaMalloc(&g.d_threads.paths.rays.d, n_threads * sizeof(float3)));
      CK(cudaMalloc(&g.d_threads.paths.throughputs, n_threads * sizeof(float4)));
      CK(cudaMalloc(&g.d_threads.image_ids, n_threads * sizeof(uint)));
      g.n_threads = n_threads;
    }
    uint zero = 0;
    CK(cudaMemcpyToSymbol(StreamingVolPTsk_kernel::d_paths_head_global, &zero, sizeof zero));
    CK(cudaMemcpyToSymbol(StreamingVolPTsk_kernel::c_seed, &seed, sizeof seed));
    void* args[] = {&g.d_threads, &g.d_out, &g.scene};
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernel(k, dim3(grid), dim3(block), args, 0, 0));
    CK(cudaEventRecord(e1));
    *grid_out = grid;
#else
    snprintf(g_err, sizeof g_err, "built without the streamingSK kernel");
    return 1;
#endif
  } else {
    auto k = RegenerationVolPTsk_kernel::d_render_single_thread_regeneration<RefScene>;
    CK(cudaOccupancyMaxPotentialBlockSize(&min_grid, &block, k, 0, 0));
    uint zero = 0;
    CK(cudaMemcpyToSymbol(RegenerationVolPTsk_kernel::paths_head_global, &zero, sizeof zero));
    CK(cudaMemcpyToSymbol(RegenerationVolPTsk_kernel::seed, &seed, sizeof seed));
    CK(cudaEventRecord(e0));
    k<<<min_grid, block>>>(g.d_out, g.scene);
    CK(cudaEventRecord(e1));
    *grid_out = min_grid;
  }
  *block_out = block;
  CK(cudaGetLastError());
  CK(cudaEventSynchronize(e1));
  CK(cudaEventElapsedTime(ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (host_out)
    CK(cudaMemcpy(host_out, g.d_out, g.out_pixels * sizeof(float4), cudaMemcpyDeviceToHost));
  return 0;
}

/* cuRAND device XORWOW, exactly as Rng.h:22,26 uses it: n words / uniforms per seed */
__global__ void k_curand_kat(const int* seeds, int n_seeds, int n, unsigned* words, float* uni) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  curandState s;
  curand_init(seeds[i], 0, 0, &s);
  for (int k = 0; k < n; ++k) words[i * n + k] = curand(&s);
  Rng r(seeds[i]);
  for (int k = 0; k < n; ++k) uni[i * n + k] = r.getFloat();
}

int refgpu_curand_kat(const int* seeds, int n_seeds, int n, unsigned* words, float* uni) {
  int* d_seeds;
  unsigned* d_w;
  float* d_u;
  CK(cudaMalloc(&d_seeds, n_seeds * sizeof(int)));
  CK(cudaMalloc(&d_w, (size_t)n_seeds * n * 4));
  CK(cudaMalloc(&d_u, (size_t)n_seeds * n * 4));
  CK(cudaMemcpy(d_seeds, seeds, n_seeds * sizeof(int), cudaMemcpyHostToDevice));
  k_curand_kat<<<(n_seeds + 63) / 64, 64>>>(d_seeds, n_seeds, n, d_w, d_u);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(words, d_w, (size_t)n_seeds * n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(uni, d_u, (size_t)n_seeds * n * 4, cudaMemcpyDeviceToHost));
  cudaFree(d_seeds), cudaFree(d_w), cudaFree(d_u);
  return 0;
}

/* utilhash on the device (Utilities.cuh:157-171) for the hash KAT */
__global__ void k_hash(const unsigned* in, int n, unsigned* out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = utilhash(in[i]);
}
int refgpu_has_streaming() {
#ifdef CVR_REF_STREAMING
  return 1;
#else
  return 0;
#endif
}
int refgpu_utilhash(const unsigned* in, int n, unsigned* out) {
  unsigned *di, *dout;
  CK(cudaMalloc(&di, n * 4));
  CK(cudaMalloc(&dout, n * 4));
  CK(cudaMemcpy(di, in, n * 4, cudaMemcpyHostToDevice));
  k_hash<<<(n + 63) / 64, 64>>>(di, n, dout);
  CK(cudaMemcpy(out, dout, n * 4, cudaMemcpyDeviceToHost));
  cudaFree(di), cudaFree(dout);
  return 0;
}

}  // extern "C"
