// Scene.h -- host-side scene model of the drop-in host layer.
//
// Mirrors the reference's Scene / SceneBuilder / SceneAssembler (Scene.h:19-81),
// Camera (Camera.h:16-71) and HostMedium = HeterogeneousMedium<Volume<float4>,
// Volume<float>, HG> (Medium.h:109-116,191) without glm / CUDA types: plain vectors
// and floats, because everything below this layer is reached through the C ABI.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cvr_abi.h"

namespace cvrhost {

struct uint2 {
  uint32_t x = 0, y = 0;
};
struct float3 {
  float x = 0, y = 0, z = 0;
};

// Camera.h:16-71: resolution, fov (y derived from x), model-view matrix
class Camera {
  int res_x_, res_y_;
  float fov_x_;
  std::array<float, 12> inv_view_{};  // rows as CudaVolPath::initCamera lays them out (CudaVolPath.cpp:71-84)
  std::array<float, 2> raster_to_view_{};

  void update() { cvr_default_camera((uint32_t)res_x_, (uint32_t)res_y_, fov_x_, nullptr, raster_to_view_.data()); }

 public:
  explicit Camera(int res_x = 400, int res_y = 400, float fov_x = 0.7f) : res_x_(res_x), res_y_(res_y), fov_x_(fov_x) {
    cvr_default_camera((uint32_t)res_x, (uint32_t)res_y, fov_x, inv_view_.data(), raster_to_view_.data());
  }
  void setResolution(int x, int y) {  // Camera.h:50-53
    res_x_ = x, res_y_ = y;
    update();
  }
  uint2 getResolution() const { return {(uint32_t)res_x_, (uint32_t)res_y_}; }
  float getFovX() const { return fov_x_; }
  const float* getInvViewMatrix() const { return inv_view_.data(); }
  void setInvViewMatrix(const float m[12]) { std::copy(m, m + 12, inv_view_.begin()); }
  std::array<float, 2> getRasterToView() const { return raster_to_view_; }  // Camera.h:69-71
};

// Volume<T> (Volume.h:116-179): dense, x-fastest
template <int CH>
struct Volume {
  std::vector<float> data;
  uint32_t nx = 0, ny = 0, nz = 0;
  size_t voxels() const { return (size_t)nx * ny * nz; }
  size_t getBytes() const { return data.size() * sizeof(float); }
  void check() const {
    if (data.size() != voxels() * CH) throw std::runtime_error("Volume data size does not match grid resolution");
  }
};

struct AABB {
  float3 box_min, box_max;
};

struct HostMedium {
  AABB density_AABB;
  float scale = 1.f;
  float max_density = 1.f;
  Volume<4> albedo_volume;  // rgb + w; empty = the constant albedo below
  float albedo_const[3] = {1.f, 1.f, 1.f};
  Volume<1> density_volume;
  float hg_g = 0.f;  // Volume.h:20
};

class Scene {
  std::shared_ptr<Camera> camera_;
  HostMedium medium_;

 public:
  Scene() = default;
  Scene(std::shared_ptr<Camera> camera, HostMedium medium) : camera_(std::move(camera)), medium_(std::move(medium)) {}
  std::shared_ptr<Camera> getCamera() const { return camera_; }
  const HostMedium& getMedium() const { return medium_; }

  cvr_scene_desc desc() const {
    cvr_scene_desc d{};
    d.density = medium_.density_volume.data.data();
    d.density_dim[0] = (int32_t)medium_.density_volume.nx, d.density_dim[1] = (int32_t)medium_.density_volume.ny;
    d.density_dim[2] = (int32_t)medium_.density_volume.nz;
    d.albedo = medium_.albedo_volume.data.empty() ? nullptr : medium_.albedo_volume.data.data();
    d.albedo_dim[0] = (int32_t)medium_.albedo_volume.nx, d.albedo_dim[1] = (int32_t)medium_.albedo_volume.ny;
    d.albedo_dim[2] = (int32_t)medium_.albedo_volume.nz;
    for (int c = 0; c < 3; ++c) d.albedo_const[c] = medium_.albedo_const[c];
    d.box_min[0] = medium_.density_AABB.box_min.x, d.box_min[1] = medium_.density_AABB.box_min.y;
    d.box_min[2] = medium_.density_AABB.box_min.z;
    d.box_max[0] = medium_.density_AABB.box_max.x, d.box_max[1] = medium_.density_AABB.box_max.y;
    d.box_max[2] = medium_.density_AABB.box_max.z;
    d.scale = medium_.scale, d.max_density = medium_.max_density, d.hg_g = medium_.hg_g;
    d.ggx_alpha[0] = d.ggx_alpha[1] = 0.1f;  // Bsdf.h:18
    d.ggx_eta = 1.05f / 1.01f;               // Bsdf.h:21-22
    return d;
  }
};

class SceneBuilder {  // Scene.h:56-63
 public:
  virtual ~SceneBuilder() = default;
  virtual std::shared_ptr<Camera> getCamera() = 0;
  virtual HostMedium getMedium() = 0;
};

class SceneAssembler {  // Scene.h:65-81
  std::unique_ptr<SceneBuilder> builder;

 public:
  void setBuilder(std::unique_ptr<SceneBuilder> b) { builder = std::move(b); }
  Scene getScene() { return Scene(builder->getCamera(), builder->getMedium()); }
};

}  // namespace cvrhost
