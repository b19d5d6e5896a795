// CudaVolPath.h -- tile scheduler + progressive renderer of the drop-in host layer.
//
// Mirrors TilingConfig (Config.h:61-78), Buffer2D (Buffer.h:54-130),
// AbstractRenderer / AbstractProgressiveRenderer (AbstractRenderer.h:8-24),
// CudaVolPath<Launcher> (CudaVolPath.h:34-102, CudaVolPath.cpp) and the factory switch
// (RendererFactory.h:37-115).  The renderer drives a launcher call by call in the
// reference's order; the caller-owned tile buffer is plain cudaMalloc memory.
#pragma once
#include <cuda_runtime_api.h>

#include <cmath>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "RenderKernelLauncher.h"

namespace cvrhost {

struct TilingConfig {  // Config.h:61-78
  uint2 n_tiles{1, 1};
  uint2 tile_dim{};
  uint2 resolution{400, 400};
  explicit TilingConfig(uint2 res = {400, 400}, uint2 tiles = {1, 1}) : n_tiles(tiles), resolution(res) {
    tile_dim.x = (uint32_t)(int)std::ceil((double)(resolution.x / n_tiles.x));  // integer division: floors (Q6)
    tile_dim.y = (uint32_t)(int)std::ceil((double)(resolution.y / n_tiles.y));
  }
};

struct Buffer2D {  // Buffer.h:54-130 (float4 host image view)
  void* data = nullptr;
  size_t width = 0, width_bytes = 0, height = 0, pitch_bytes = 0;
};
inline Buffer2D make_buffer2D_float4(float* data, size_t w, size_t h) {
  return {data, w, w * 16, h, w * 16};
}

class AbstractRenderer {
 public:
  virtual ~AbstractRenderer() = default;
  virtual void render(Buffer2D buffer_out) = 0;
};

class AbstractProgressiveRenderer : public AbstractRenderer {
 public:
  virtual void initRendering() = 0;
  virtual void runIterations() = 0;
  virtual void setNIterations(uint32_t n) = 0;
  virtual bool imageComplete() = 0;
  virtual void getImage(Buffer2D buffer_out) = 0;
};

inline void cuda_ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

typedef std::vector<std::pair<std::string, std::string>> LauncherOptions;

template <class VolPathKernelLauncher>
class CudaVolPath : public AbstractProgressiveRenderer {
  TilingConfig tiling_config_;
  uint32_t iterations_;
  Scene scene_;
  uint32_t current_iteration_ = 0;
  std::vector<uint2> tiles_;
  size_t current_tile_ = 0;
  void* d_output_ = nullptr;  // tile accumulation buffer (float4)
  void* d_image_ = nullptr;   // resolved full image (float4)
  VolPathKernelLauncher kernel_launcher_;

  void initTileArray() {  // CudaVolPath.cpp:12-29
    uint32_t n = tiling_config_.n_tiles.x * tiling_config_.n_tiles.y;
    std::vector<uint32_t> org(2 * (size_t)n);
    uint32_t dim[2];
    cvr_tile_table(tiling_config_.resolution.x, tiling_config_.resolution.y, tiling_config_.n_tiles.x,
                   tiling_config_.n_tiles.y, dim, org.data());
    tiles_.resize(n);
    for (uint32_t i = 0; i < n; ++i) tiles_[i] = {org[2 * i], org[2 * i + 1]};
    current_tile_ = 0;
  }

 public:
  // `options` = key/value pairs for cvr_set_option, applied right after the launcher exists and
  // BEFORE init() / setScene() (layout, sched, rng ... select the kernel and the device layout)
  CudaVolPath(const Scene& scene, TilingConfig tiling, uint32_t iterations, int device = 0,
              const LauncherOptions& options = {})
      : tiling_config_(tiling), iterations_(iterations), scene_(scene), kernel_launcher_(device) {
    for (const auto& kv : options) kernel_launcher_.setOption(kv.first, kv.second);
    // constructor order of CudaVolPath.cpp:31-59
    initTileArray();
    auto rtv = scene_.getCamera()->getRasterToView();
    kernel_launcher_.copyRasterToView(rtv[0], rtv[1]);
    kernel_launcher_.setResolution(tiling_config_.tile_dim);
    kernel_launcher_.copyPixelIndexRange((float)tiling_config_.resolution.x, (float)tiling_config_.resolution.y);
    kernel_launcher_.init();
    size_t tile_px = (size_t)tiling_config_.tile_dim.x * tiling_config_.tile_dim.y;
    cuda_ck(cudaMalloc(&d_output_, tile_px * 16), "cudaMalloc tile");  // allocateDeviceMemory :211-232
    cuda_ck(cudaMalloc(&d_image_, (size_t)tiling_config_.resolution.x * tiling_config_.resolution.y * 16), "cudaMalloc image");
    kernel_launcher_.setOutputPtr(d_output_);
    kernel_launcher_.allocateDeviceMemory();
    kernel_launcher_.setScene(scene_);  // initDeviceScene :87-115
  }
  ~CudaVolPath() override {
    cudaDeviceSynchronize();
    cudaFree(d_output_);
    cudaFree(d_image_);
  }
  VolPathKernelLauncher& launcher() { return kernel_launcher_; }

  void setNIterations(uint32_t n) override {
    iterations_ = n;
    kernel_launcher_.setNIterations(n);
  }
  void initRendering() override {  // initCamera + initRenderState (:66-85, :202-208)
    kernel_launcher_.copyInvViewMatrix(scene_.getCamera()->getInvViewMatrix(), 48);
    current_iteration_ = 0;
    size_t tile_px = (size_t)tiling_config_.tile_dim.x * tiling_config_.tile_dim.y;
    cuda_ck(cudaMemset(d_output_, 0, tile_px * 16), "cudaMemset");
    current_tile_ = 0;
  }
  bool imageComplete() override { return current_tile_ == tiles_.size(); }
  void runIterations() override {  // :248-280
    if (current_tile_ == tiles_.size()) current_tile_ = 0;
    if (current_tile_ == 0) current_iteration_ += kernel_launcher_.getNIterations();
    kernel_launcher_.copyOffset(tiles_[current_tile_]);
    kernel_launcher_.launchRender();
    ++current_tile_;
  }
  void getImage(Buffer2D out) override {  // :282-295 + prepareForNextIterations :188-200
    uint2 tile_start = tiles_[current_tile_ - 1];
    uint2 td = tiling_config_.tile_dim, full = tiling_config_.resolution;
    // intended transfer semantics (the reference's delegate is broken at HEAD, Q13): every
    // float of the tile divided by current_iteration_, copied at the tile origin
    kernel_launcher_.resolveTile(d_output_, td, d_image_, full, tile_start, (float)current_iteration_);
    kernel_launcher_.reset();  // sync + seed advance
    size_t off = ((size_t)tile_start.y * full.x + tile_start.x) * 16;
    cuda_ck(cudaMemcpy2D((char*)out.data + (size_t)tile_start.y * out.pitch_bytes + (size_t)tile_start.x * 16,
                         out.pitch_bytes, (char*)d_image_ + off, (size_t)full.x * 16, (size_t)td.x * 16, td.y,
                         cudaMemcpyDeviceToHost), "cudaMemcpy2D");
    if (tiles_.size() != 1) {  // with one tile the buffer keeps accumulating
      cuda_ck(cudaMemset(d_output_, 0, (size_t)td.x * td.y * 16), "cudaMemset");
    }
  }
  void render(Buffer2D out) override {  // :338-347
    setNIterations(iterations_);
    initRendering();
    while (!imageComplete()) {
      runIterations();
      getImage(out);
    }
  }
};

// Multi-GPU form of CudaVolPath::render: one launcher per device of this process behind the C ABI's
// device group (cvr_group_*: one host thread per device, the scene replicated, work split by tile
// and/or sample index, the framebuffers summed into the first device's with ONE ncclReduce).  The
// image equals the single-GPU render of the same kernel name and seed.
class GroupVolPath : public AbstractRenderer {
  cvr_group_handle g_ = nullptr;
  TilingConfig tiling_config_;
  uint32_t iterations_;
  Scene scene_;
  int shard_mode_;

  void ck(int rc, const char* what) const {
    if (rc) throw std::runtime_error(std::string(what) + ": " + cvr_group_last_error(g_));
  }

 public:
  GroupVolPath(const std::string& kernel, const Scene& scene, TilingConfig tiling, uint32_t iterations, int n_devices,
               int shard_mode, const LauncherOptions& options = {})
      : tiling_config_(tiling), iterations_(iterations), scene_(scene), shard_mode_(shard_mode) {
    if (cvr_group_create(kernel.c_str(), nullptr, n_devices, &g_))
      throw std::runtime_error(std::string("cvr_group_create: ") + cvr_group_last_error(nullptr));
    try {
      for (const auto& kv : options) ck(cvr_group_set_option(g_, kv.first.c_str(), kv.second.c_str()), "setOption");
      cvr_scene_desc d = scene_.desc();
      ck(cvr_group_set_scene(g_, &d), "setScene");
    } catch (...) {
      cvr_group_destroy(g_);
      throw;
    }
  }
  ~GroupVolPath() override { cvr_group_destroy(g_); }
  GroupVolPath(const GroupVolPath&) = delete;
  GroupVolPath& operator=(const GroupVolPath&) = delete;
  cvr_group_handle group() const { return g_; }

  void render(Buffer2D out) override {
    if (out.pitch_bytes != (size_t)tiling_config_.resolution.x * 16)
      throw std::runtime_error("GroupVolPath::render: the output image must be dense float4 rows");
    cvr_render_desc r{};
    r.res_x = tiling_config_.resolution.x, r.res_y = tiling_config_.resolution.y;
    r.n_tiles_x = tiling_config_.n_tiles.x, r.n_tiles_y = tiling_config_.n_tiles.y;
    r.iterations = iterations_;
    r.fov_x = scene_.getCamera()->getFovX();
    auto rtv = scene_.getCamera()->getRasterToView();
    r.inv_view = scene_.getCamera()->getInvViewMatrix();
    r.raster_to_view = rtv.data();
    r.fuse_tiles = 1;
    ck(cvr_group_render_image(g_, &r, shard_mode_, (float*)out.data, nullptr), "render");
  }
};

// RendererFactory::createRenderer (RendererFactory.h:13-22,37-115): kernel name -> renderer
inline std::unique_ptr<AbstractProgressiveRenderer> createRenderer(const std::string& kernel, const Scene& scene,
                                                                   TilingConfig tiling, uint32_t iterations,
                                                                   int device = 0, const LauncherOptions& options = {}) {
  if (kernel == "naiveSK") return std::make_unique<CudaVolPath<NaiveVolPTsk>>(scene, tiling, iterations, device, options);
  if (kernel == "regenerationSK") return std::make_unique<CudaVolPath<RegenerationVolPTsk>>(scene, tiling, iterations, device, options);
  if (kernel == "streamingSK") return std::make_unique<CudaVolPath<StreamingVolPTsk>>(scene, tiling, iterations, device, options);
  if (kernel == "streamingMK") return std::make_unique<CudaVolPath<StreamingVolPTmk>>(scene, tiling, iterations, device, options);
  if (kernel == "sortingSK") return std::make_unique<CudaVolPath<SortingVolPTsk>>(scene, tiling, iterations, device, options);
  throw std::runtime_error("kernel '" + kernel + "' is not available in this build (naiveSK | regenerationSK | streamingSK | streamingMK | sortingSK)");
}

}  // namespace cvrhost
