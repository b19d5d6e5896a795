// cvr_scene_file.cpp -- C entry points over the C++ scene loaders of the host layer
// (host/SceneBuilders.h: RawSceneBuilder.h:35-140, XmlSceneBuilder.h:39-266, VDBSceneBuilder.h:40-80 and the
// procedural stand-ins), so that every binding -- the ctypes layer of this repo, a cgo / JNI stub -- loads scenes
// through the SAME code as cvr_render instead of restating the loaders' quirks.  Host-only.
#include <cstring>
#include <string>

#include "../../include/cvr_abi.h"
#include "../host/SceneBuilders.h"

struct cvr_scene_file {
  cvrhost::Scene scene;
  std::string type;
};

static thread_local std::string g_scene_file_error;

static int scene_file_fail(const std::string& m) {
  g_scene_file_error = m;
  return 1;
}

extern "C" {

const char* cvr_scene_file_last_error(void) { return g_scene_file_error.c_str(); }

int cvr_scene_file_load(const char* path, const char* type, cvr_scene_file_handle* out) {
  if (!path || !out) return scene_file_fail("cvr_scene_file_load: null argument");
  *out = nullptr;
  try {
    std::string resolved;
    cvrhost::SceneAssembler assembler;
    assembler.setBuilder(cvrhost::makeSceneBuilder(path, type && *type ? type : "Auto", &resolved));
    *out = new cvr_scene_file{assembler.getScene(), resolved};
    return 0;
  } catch (const std::exception& e) {
    return scene_file_fail(e.what());
  }
}

int cvr_scene_file_close(cvr_scene_file_handle f) {
  delete f;
  return 0;
}

int cvr_scene_file_info(cvr_scene_file_handle f, cvr_scene_file_info_t* info) {
  if (!f || !info) return scene_file_fail("cvr_scene_file_info: null argument");
  std::memset(info, 0, sizeof(*info));
  info->scene = f->scene.desc();  // pointers into the handle's volumes: valid until cvr_scene_file_close
  const auto cam = f->scene.getCamera();
  const cvrhost::uint2 res = cam->getResolution();
  info->resolution[0] = res.x, info->resolution[1] = res.y;
  info->fov_x = cam->getFovX();
  std::memcpy(info->inv_view, cam->getInvViewMatrix(), sizeof(info->inv_view));
  const auto rtv = cam->getRasterToView();
  info->raster_to_view[0] = rtv[0], info->raster_to_view[1] = rtv[1];
  std::strncpy(info->type, f->type.c_str(), sizeof(info->type) - 1);
  return 0;
}

}  // extern "C"
