/* B200VolPTKernelLauncher.h -- the reference-side binding of libcvr_b200.so.
 *
 * A maintainer of CudaVolumeRenderer drops this file into implementation/src/: it is a launcher class
 * in the reference's own plugin hierarchy (RenderKernelLauncher / VolPTKernelLauncher<DeviceScene>,
 * RenderKernelLauncher.h:20-73) whose members forward to the C ABI of include/cvr_abi.h, so that
 * CudaVolPath<B200VolPTsk<DeviceScene>> (CudaVolPath.h:34-102) drives the B200 kernels through the call
 * sequence it already has (CudaVolPath.cpp:41-55,84,114,190,229-263,330).
 *
 * It needs the reference's headers and is therefore NOT part of the product build;
 * oracle/ref_binding_check.cpp compiles it against /root/reference and runs that call sequence
 * (oracle/Makefile target `ref_binding`, tests/test_oracle.py::test_reference_side_binding_compiles_and_links, tests/test_gpu_parity.py::test_reference_tile_driver_runs_on_the_documented_binding).
 */
#pragma once
#include <cstdio>
#include <cstdlib>

#include "RenderKernelLauncher.h" /* the reference's */
#include "cvr_abi.h"

template <class DEVICE_SCENE>
class B200VolPTsk : public VolPTKernelLauncher<DEVICE_SCENE> {
  using uint = unsigned int;
  cvr_handle h_ = nullptr;
  static void ck(cvr_handle h, int rc) { /* reference convention: print + exit (Debug.h:21-37) */
    if (rc) {
      fprintf(stderr, "cvr: %s\n", cvr_last_error(h));
      exit(EXIT_FAILURE);
    }
  }

 public:
  explicit B200VolPTsk(const char* kernel = "regenerationSK") {
    int dev = 0;
    cudaGetDevice(&dev);
    ck(nullptr, cvr_create(kernel, dev, &h_));
    /* the reference works on the default stream (cudaMemset of the tile buffer, CudaVolPath.cpp:196-207):
     * run the kernels on the legacy stream so that the stream itself orders them */
    ck(h_, cvr_set_stream(h_, (void*)cudaStreamLegacy));
  }
  B200VolPTsk(const B200VolPTsk&) = delete;
  B200VolPTsk& operator=(const B200VolPTsk&) = delete;
  ~B200VolPTsk() { cvr_destroy(h_); }
  cvr_handle handle() const { return h_; }

  /* RenderKernelLauncher.h:31-51 (non-virtual members are hidden: CudaVolPath names the concrete type) */
  void setOutputPtr(float4* d_output) {
    this->d_output_ = d_output;
    ck(h_, cvr_set_output(h_, d_output));
  }
  void setResolution(uint2 r) {
    this->resolution_ = r;
    ck(h_, cvr_set_resolution(h_, r.x, r.y));
  }
  void setCudaConfig(CudaConfig c) { this->cuda_config_ = c; } /* launch shape is the library's (cvr_get_launch_shape) */
  void init() override { ck(h_, cvr_init(h_)); }
  void allocateDeviceMemory() override { ck(h_, cvr_allocate(h_)); }
  void launchRender() override { ck(h_, cvr_launch_render(h_)); }
  void reset() override { ck(h_, cvr_reset(h_)); } /* sync + seed advance */
  void releaseDeviceMemory() override { ck(h_, cvr_release(h_)); }
  void copyInvViewMatrix(float* m, size_t) { ck(h_, cvr_set_inv_view_matrix(h_, m)); }
  void copyRasterToView(float2 v) { ck(h_, cvr_set_raster_to_view(h_, v.x, v.y)); }
  void copyPixelIndexRange(float2 v) { ck(h_, cvr_set_pixel_index_range(h_, v.x, v.y)); }
  void copyOffset(uint2 o) { ck(h_, cvr_set_offset(h_, o.x, o.y)); }
  /* VolPTKernelLauncher (RenderKernelLauncher.h:54-73) */
  void setNIterations(uint n) override {
    this->n_iterations_ = n;
    ck(h_, cvr_set_iterations(h_, n));
  }
  /* replaces setScene + CudaVolPath::initDeviceScene (CudaVolPath.cpp:87-115): the host volumes go straight
   * to the library, which builds its own device layout */
  void setHostScene(const Scene& scene) {
    const auto& m = scene.getMedium();
    cvr_scene_desc d{};
    d.density = m.density_volume.getVolumeData();
    d.density_dim[0] = m.density_volume.grid_resolution.x;
    d.density_dim[1] = m.density_volume.grid_resolution.y;
    d.density_dim[2] = m.density_volume.grid_resolution.z;
    d.albedo = (const float*)m.albedo_volume.getVolumeData();
    d.albedo_dim[0] = m.albedo_volume.grid_resolution.x;
    d.albedo_dim[1] = m.albedo_volume.grid_resolution.y;
    d.albedo_dim[2] = m.albedo_volume.grid_resolution.z;
    d.box_min[0] = m.density_AABB.box_min.x, d.box_min[1] = m.density_AABB.box_min.y, d.box_min[2] = m.density_AABB.box_min.z;
    d.box_max[0] = m.density_AABB.box_max.x, d.box_max[1] = m.density_AABB.box_max.y, d.box_max[2] = m.density_AABB.box_max.z;
    d.scale = m.scale;
    d.max_density = m.max_density;
    d.hg_g = m.phase.g;                             /* Volume.h:20: g = 0 unless a builder sets it (Q5) */
    d.ggx_alpha[0] = d.ggx_alpha[1] = 0.1f;         /* Bsdf.h:18-22 */
    d.ggx_eta = 1.05f / 1.01f;
    ck(h_, cvr_set_scene(h_, &d));
  }
};
