/*
 * ref_binding_check.cpp -- builds oracle/_ref/ref_binding_check: the REFERENCE's own tile driver
 * CudaVolPath<Launcher> (CudaVolPath.h:34-102, CudaVolPath.cpp, compiled from /root/reference where it
 * lies) instantiated over the launcher class INTEGRATION.md documents (include/B200VolPTKernelLauncher.h),
 * i.e. the reference's control flow driving libcvr_b200.so through the C ABI.
 *
 * TEST INFRASTRUCTURE ONLY.  It proves that the documented binding satisfies the plugin interface
 * (RenderKernelLauncher.h:20-73): it compiles and links against the reference's headers, and -- on a GPU
 * box -- CudaVolPath::render() produces the image cvr_render_image() produces for the same scene.
 * This TU holds no reference code, only
 *  - the three wiring edits of INTEGRATION.md section 1: the launcher type, the explicit specialisation of
 *    initDeviceScene (host volumes go to the library instead of two cudaArray textures) and the
 *    instantiation list (CUDAVOLPATH_TEMPLATES, Defines.h:112, re-pointed at the one launcher);
 *  - what RenderKernelLauncher.cu supplies to the base class (the out-of-line setNIterations);
 *  - an output delegate that resolves with cvr_resolve_tile;
 *  - a small scene built with the reference's host classes (Volume, HostMedium, Camera, Scene, Config).
 * glm is not in this image: oracle/glm_standin/ provides the few names Camera.h uses.
 */
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

/* Host translation unit, like the reference's own CudaVolPath.cpp (compiled by the host compiler there too):
 * without __CUDACC__ the CUDA qualifiers are empty macros.  Names that only never-instantiated reference
 * templates mention (the gradient medium, Medium.h:30-106) still need a declaration for two-phase lookup. */
inline float norm3df(float a, float b, float c) { return std::sqrt(a * a + b * b + c * c); }
#pragma push_macro("__forceinline__")
#undef __forceinline__
#define __forceinline__

#include "B200VolPTKernelLauncher.h" /* includes the reference's RenderKernelLauncher.h */
#include "CudaVolPath.h"

typedef SimpleVolumeDeviceScene<DeviceMedium, GGX> RefScene;
static const char* g_kernel = "regenerationSK";
/* CudaVolPath default-constructs its launcher (CudaVolPath.h:70): the kernel name comes from where
 * RendererFactory would pick the type (RendererFactory.h:37-115) */
struct Launcher : public B200VolPTsk<RefScene> {
  Launcher() : B200VolPTsk<RefScene>(g_kernel) {}
};

/* edit 1 of 3: the instantiation list names the one launcher */
#undef CUDAVOLPATH_TEMPLATES
#define CUDAVOLPATH_TEMPLATES
/* edit 2 of 3: initDeviceScene hands the host volumes to the library (declared before the generic
 * definition in CudaVolPath.cpp is seen) */
template <>
void CudaVolPath<Launcher>::initDeviceScene();

#include "CudaVolPath.cpp"

#pragma pop_macro("__forceinline__")

template <>
void CudaVolPath<Launcher>::initDeviceScene() {
  kernel_launcher_.setHostScene(scene_);
}
/* edit 3 of 3 */
template class CudaVolPath<Launcher>;

/* RenderKernelLauncher.cu:122-127 equivalent for the base class' vtable (the binding overrides it) */
template <class DEVICE_SCENE>
void VolPTKernelLauncher<DEVICE_SCENE>::setNIterations(uint n_iterations) {
  n_iterations_ = n_iterations;
}
template class VolPTKernelLauncher<RefScene>;

/* the transfer delegate of the non-interactive path (ImageBufferTransfer.cu:6-18,61-78) on cvr_resolve_tile */
struct ResolveDelegate : public Buffer2DTransferDelegate<UtilityFunctors::Scale> {
  cvr_handle h = nullptr;
  uint2 full{};
  ResolveDelegate(uint2 f) : full(f) {
    if (cvr_create("naiveSK", 0, &h)) exit(3);
    cvr_set_stream(h, (void*)cudaStreamLegacy);
  }
  ~ResolveDelegate() { cvr_destroy(h); }
  void transfer(Buffer2D in, Buffer2D out, uint2 off, UtilityFunctors::Scale s) override {
    if (cvr_resolve_tile(h, in.data, (uint32_t)in.width, (uint32_t)in.height, out.data, full.x, full.y, off.x, off.y, s.scale)) {
      fprintf(stderr, "cvr_resolve_tile: %s\n", cvr_last_error(h));
      exit(4);
    }
  }
};

static float voxel(int x, int y, int z, int n) {
  const float fx = (x + 0.5f) / n - 0.5f, fy = (y + 0.5f) / n - 0.5f, fz = (z + 0.5f) / n - 0.5f;
  const float r = std::sqrt(fx * fx + fy * fy + fz * fz);
  const float v = 0.5f + 0.5f * std::sin(17.0f * fx) * std::cos(13.0f * fy + 5.0f * fz);
  return r < 0.45f ? v : 0.0f;
}

int main(int argc, char** argv) {
  const char* kernel = argc > 1 ? argv[1] : "naiveSK";
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    printf("ref_binding_check: built and linked; no CUDA device, nothing run\n");
    return 0;
  }
  const unsigned n = 24, res = 96, spp = 8, tiles = 2;
  std::vector<float> dens(n * n * n);
  std::vector<float4> alb(n * n * n);
  float vmax = 0.f;
  for (unsigned z = 0; z < n; ++z)
    for (unsigned y = 0; y < n; ++y)
      for (unsigned x = 0; x < n; ++x) {
        const float v = voxel(x, y, z, n);
        dens[x + n * (y + n * z)] = v;
        alb[x + n * (y + n * z)] = make_float4(0.9f, 0.5f + 0.4f * v, 0.3f, 1.0f);
        vmax = std::fmax(vmax, v);
      }
  std::vector<float> dens_copy = dens;
  std::vector<float4> alb_copy = alb;

  /* the scene in the reference's own host classes */
  HostMedium medium;
  medium.density_AABB = AABB(make_float3(-10, -10, -10), make_float3(10, 10, 10));
  medium.scale = 0.8f;
  medium.max_density = vmax;
  medium.density_volume = Volume<float>(std::move(dens), make_uint3(n, n, n));
  medium.albedo_volume = Volume<float4>(std::move(alb), make_uint3(n, n, n));
  auto camera = std::make_shared<Camera>(res, res, 15.0f);
  Scene scene(camera, std::shared_ptr<AbstractGeometry>(), medium);
  Config config(scene);
  config.path_tracing_config = PathTracingConfig(spp);
  config.tiling_config = TilingConfig(make_uint2(res, res), make_uint2(tiles, tiles));

  float4* d_image = nullptr;
  cudaMalloc(&d_image, sizeof(float4) * res * res);
  cudaMemset(d_image, 0, sizeof(float4) * res * res);
  std::vector<float4> via_reference_driver(res * res), direct(res * res);
  {
    g_kernel = kernel;
    CudaVolPath<Launcher> renderer(config, std::make_unique<ResolveDelegate>(make_uint2(res, res)));
    renderer.render(make_buffer2D<float4>(d_image, res, res));
    cudaDeviceSynchronize();
    cudaMemcpy(via_reference_driver.data(), d_image, sizeof(float4) * res * res, cudaMemcpyDeviceToHost);
  }

  /* the same render through the C ABI alone */
  cvr_handle h = nullptr;
  if (cvr_create(kernel, 0, &h)) return 5;
  cvr_scene_desc d{};
  d.density = dens_copy.data(), d.albedo = (const float*)alb_copy.data();
  for (int i = 0; i < 3; ++i) d.density_dim[i] = d.albedo_dim[i] = n, d.box_min[i] = -10, d.box_max[i] = 10;
  d.scale = 0.8f, d.max_density = vmax, d.hg_g = 0.f, d.ggx_alpha[0] = d.ggx_alpha[1] = 0.1f, d.ggx_eta = 1.05f / 1.01f;
  const float2 r2v = camera->getRasterToView();
  const float raster_to_view[2] = {r2v.x, r2v.y};
  cvr_render_desc rd{};
  rd.res_x = rd.res_y = res, rd.n_tiles_x = rd.n_tiles_y = tiles, rd.iterations = spp;
  rd.raster_to_view = raster_to_view; /* inv_view NULL = the default camera (Camera.h:25-37), what `camera` holds */
  std::vector<float4> host(res * res);
  int rc = cvr_set_scene(h, &d) || cvr_render_image(h, &rd, (float*)host.data(), d_image);
  if (rc) {
    fprintf(stderr, "direct: %s\n", cvr_last_error(h));
    return 6;
  }
  cvr_sync(h);
  cudaMemcpy(direct.data(), d_image, sizeof(float4) * res * res, cudaMemcpyDeviceToHost);
  cvr_destroy(h);

  /* same seeds, same tile order, same kernels: the two images differ by the order of the fp32 atomic adds only */
  double sum_a = 0, sum_b = 0, max_diff = 0;
  size_t hit = 0;
  for (size_t i = 0; i < via_reference_driver.size(); ++i) {
    const float4 a = via_reference_driver[i], b = direct[i];
    sum_a += a.x + a.y + a.z, sum_b += b.x + b.y + b.z;
    max_diff = std::fmax(max_diff, std::fmax(std::fabs(a.x - b.x), std::fmax(std::fabs(a.y - b.y), std::fabs(a.z - b.z))));
    hit += a.x != 1.0f;
  }
  printf("ref_binding_check: reference CudaVolPath<B200VolPTsk> mean %.6f, cvr_render_image mean %.6f, max |diff| %.3g, "
         "%zu of %zu pixels see the volume\n",
         sum_a / (3.0 * res * res), sum_b / (3.0 * res * res), max_diff, hit, via_reference_driver.size());
  const bool ok = max_diff <= 1e-4 && hit > via_reference_driver.size() / 20 && std::fabs(sum_a - sum_b) <= 1e-5 * sum_b;
  printf(ok ? "ref_binding_check: OK\n" : "ref_binding_check: MISMATCH\n");
  return ok ? 0 : 1;
}
