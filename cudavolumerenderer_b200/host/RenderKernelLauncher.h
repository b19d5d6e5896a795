// RenderKernelLauncher.h -- C++ mirror of the reference's launcher plugin surface
// (RenderKernelLauncher.h:20-73 and the launcher classes :75-82, :104-113, :139-151)
// implemented as thin calls into the C ABI (include/cvr_abi.h).  Same method names and
// argument meaning; CUDA/config errors surface as std::runtime_error carrying the
// library's message (the reference prints and exit()s, Debug.h:21-37).
#pragma once
#include <stdexcept>
#include <string>

#include "Scene.h"

namespace cvrhost {

class RenderKernelLauncher {
 protected:
  cvr_handle h_ = nullptr;
  uint2 resolution_{};
  void* d_output_ = nullptr;

  void ck(int rc, const char* what) const {
    if (rc) throw std::runtime_error(std::string(what) + ": " + cvr_last_error(h_));
  }

 public:
  RenderKernelLauncher(const char* kernel, int device) {
    if (cvr_create(kernel, device, &h_)) throw std::runtime_error(std::string("cvr_create: ") + cvr_last_error(nullptr));
    // The reference launches on the default stream and its caller clears / copies the tile buffer
    // with plain cudaMemset / cudaMemcpy2D (CudaVolPath.cpp:196-207): keep that ordering.  The
    // handle's own stream is non-blocking and would NOT order against those calls (a memset could
    // still be zeroing the tile while the next launch accumulates into it).
    if (cvr_set_stream(h_, (void*)0x1 /* cudaStreamLegacy */)) {
      std::string msg = std::string("cvr_set_stream: ") + cvr_last_error(h_);
      cvr_destroy(h_);
      throw std::runtime_error(msg);
    }
  }
  virtual ~RenderKernelLauncher() { cvr_destroy(h_); }
  RenderKernelLauncher(const RenderKernelLauncher&) = delete;
  RenderKernelLauncher& operator=(const RenderKernelLauncher&) = delete;

  cvr_handle handle() const { return h_; }
  void setOption(const std::string& key, const std::string& value) { ck(cvr_set_option(h_, key.c_str(), value.c_str()), "setOption"); }

  void setOutputPtr(void* d_output) { d_output_ = d_output, ck(cvr_set_output(h_, d_output), "setOutputPtr"); }
  void setResolution(uint2 r) { resolution_ = r, ck(cvr_set_resolution(h_, r.x, r.y), "setResolution"); }
  virtual void allocateDeviceMemory() { ck(cvr_allocate(h_), "allocateDeviceMemory"); }
  virtual void init() { ck(cvr_init(h_), "init"); }
  virtual void launchRender() { ck(cvr_launch_render(h_), "launchRender"); }
  virtual void reset() { ck(cvr_reset(h_), "reset"); }
  virtual void releaseDeviceMemory() { ck(cvr_release(h_), "releaseDeviceMemory"); }
  void copyInvViewMatrix(const float* m, size_t /*size_of_mat*/) { ck(cvr_set_inv_view_matrix(h_, m), "copyInvViewMatrix"); }
  void copyRasterToView(float x, float y) { ck(cvr_set_raster_to_view(h_, x, y), "copyRasterToView"); }
  void copyPixelIndexRange(float w, float h) { ck(cvr_set_pixel_index_range(h_, w, h), "copyPixelIndexRange"); }
  void copyOffset(uint2 o) { ck(cvr_set_offset(h_, o.x, o.y), "copyOffset"); }
  void resolveTile(const void* d_tile, uint2 tile, void* d_image, uint2 full, uint2 off, float scale) {
    ck(cvr_resolve_tile(h_, d_tile, tile.x, tile.y, d_image, full.x, full.y, off.x, off.y, scale), "resolveTile");
  }
  cvr_counters counters() {
    cvr_counters c{};
    ck(cvr_get_counters(h_, &c), "counters");
    return c;
  }
};

class VolPTKernelLauncher : public RenderKernelLauncher {
 protected:
  uint32_t n_iterations_{1};

 public:
  using RenderKernelLauncher::RenderKernelLauncher;
  virtual void setNIterations(uint32_t n) { n_iterations_ = n, ck(cvr_set_iterations(h_, n), "setNIterations"); }
  uint32_t getNIterations() const { return n_iterations_; }
  void setScene(const Scene& scene) {
    cvr_scene_desc d = scene.desc();
    ck(cvr_set_scene(h_, &d), "setScene");
  }
};

struct NaiveVolPTsk : VolPTKernelLauncher {
  explicit NaiveVolPTsk(int device = 0) : VolPTKernelLauncher("naiveSK", device) {}
};
struct RegenerationVolPTsk : VolPTKernelLauncher {
  explicit RegenerationVolPTsk(int device = 0) : VolPTKernelLauncher("regenerationSK", device) {}
};
struct StreamingVolPTsk : VolPTKernelLauncher {
  explicit StreamingVolPTsk(int device = 0) : VolPTKernelLauncher("streamingSK", device) {}
};
struct StreamingVolPTmk : VolPTKernelLauncher {  // RenderKernelLauncher.h:115-137
  explicit StreamingVolPTmk(int device = 0) : VolPTKernelLauncher("streamingMK", device) {}
};
struct SortingVolPTsk : VolPTKernelLauncher {  // RenderKernelLauncher.h:153-170
  explicit SortingVolPTsk(int device = 0) : VolPTKernelLauncher("sortingSK", device) {}
};

}  // namespace cvrhost
