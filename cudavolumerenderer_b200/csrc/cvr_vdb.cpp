// cvr_vdb.cpp -- C entry points over host/VdbReader.h (the OpenVDB-free .vdb reader that
// replaces the reference's vdb_adapter, implementation/vdb_adapter/VDBAdapter.{h,cpp}).
// Host-only code; lives in libcvr_b200.so so that bindings (ctypes, cgo, ...) reach the
// loader through the same library as the kernels.
#include <cstring>
#include <string>

#include "../../include/cvr_abi.h"
#include "../host/VdbReader.h"

struct cvr_vdb_file {
  cvrvdb::File file;
};

static thread_local std::string g_vdb_error;

static int vdb_fail(const std::string& m) {
  g_vdb_error = m;
  return 1;
}

extern "C" {

const char* cvr_vdb_last_error(void) { return g_vdb_error.c_str(); }

int cvr_vdb_open(const char* path, cvr_vdb_handle* out) {
  if (!path || !out) return vdb_fail("cvr_vdb_open: null argument");
  *out = nullptr;
  try {
    auto* h = new cvr_vdb_file{cvrvdb::File::load(path)};
    *out = h;
    return 0;
  } catch (const std::exception& e) {
    return vdb_fail(e.what());
  }
}

int cvr_vdb_close(cvr_vdb_handle h) {
  delete h;
  return 0;
}

int cvr_vdb_grid_count(cvr_vdb_handle h, int32_t* n) {
  if (!h || !n) return vdb_fail("cvr_vdb_grid_count: null argument");
  *n = (int32_t)h->file.grids.size();
  return 0;
}

int cvr_vdb_grid_info(cvr_vdb_handle h, int32_t index, cvr_vdb_grid_info_t* info) {
  if (!h || !info) return vdb_fail("cvr_vdb_grid_info: null argument");
  if (index < 0 || (size_t)index >= h->file.grids.size()) return vdb_fail("cvr_vdb_grid_info: index out of range");
  const cvrvdb::Grid& g = h->file.grids[(size_t)index];
  std::memset(info, 0, sizeof *info);
  std::strncpy(info->name, g.name.c_str(), sizeof info->name - 1);
  std::strncpy(info->type, g.type.c_str(), sizeof info->type - 1);
  info->channels = g.channels;
  info->compression = g.compression;
  info->file_version = h->file.file_version;
  for (int a = 0; a < 3; ++a) {
    info->bbox_min[a] = g.bbox_min[a], info->bbox_max[a] = g.bbox_max[a];
    info->dim[a] = g.dim(a);
    info->background[a] = g.background[a];
  }
  info->active_voxels = g.active_voxels;
  info->leaf_count = g.leaves.size();
  info->active_tiles = g.tiles.size();
  return 0;
}

int cvr_vdb_grid_meta(cvr_vdb_handle h, const char* grid, const char* key, char* value, size_t cap) {
  if (!h || !key || !value || !cap) return vdb_fail("cvr_vdb_grid_meta: null argument");
  const std::map<std::string, std::string>* m = &h->file.meta;
  if (grid && *grid) {
    const cvrvdb::Grid* g = h->file.find(grid);
    if (!g) return vdb_fail(std::string("VDB file does not contain a ") + grid + " grid");
    m = &g->meta;
  }
  auto it = m->find(key);
  if (it == m->end()) return vdb_fail(std::string("no metadata '") + key + "'");
  std::strncpy(value, it->second.c_str(), cap - 1);
  value[cap - 1] = 0;
  return 0;
}

int cvr_vdb_densify(cvr_vdb_handle h, const char* grid, int32_t out_channels, const float* inactive, float* out,
                    uint64_t out_floats) {
  if (!h || !grid || !out) return vdb_fail("cvr_vdb_densify: null argument");
  const cvrvdb::Grid* g = h->file.find(grid);
  // message of VDBAdapter::loadVDBFile (VDBAdapter.cpp:21-37)
  if (!g || g->channels == 0) return vdb_fail(std::string("VDB file does not contain a") + (std::strcmp(grid, "albedo") ? " " : "n ") + grid + " grid");
  if (out_channels < g->channels || out_channels > 4) return vdb_fail("cvr_vdb_densify: bad channel count");
  const uint64_t need = (uint64_t)g->dim(0) * g->dim(1) * g->dim(2) * (uint64_t)out_channels;
  if (out_floats != need) return vdb_fail("cvr_vdb_densify: output size does not match the grid's active bounding box");
  const float zero[3] = {0.f, 0.f, 0.f};
  g->densify(out, out_channels, inactive ? inactive : zero);
  return 0;
}

int cvr_vdb_leaves(cvr_vdb_handle h, const char* grid, uint64_t first, uint64_t count, int32_t* origins_xyz,
                   uint64_t* masks8, float* values) {
  if (!h || !grid) return vdb_fail("cvr_vdb_leaves: null argument");
  const cvrvdb::Grid* g = h->file.find(grid);
  if (!g || g->channels == 0) return vdb_fail(std::string("VDB file does not contain a ") + grid + " grid");
  if (first + count > g->leaves.size()) return vdb_fail("cvr_vdb_leaves: range out of bounds");
  for (uint64_t i = 0; i < count; ++i) {
    const cvrvdb::Leaf& L = g->leaves[first + i];
    if (origins_xyz) std::memcpy(origins_xyz + 3 * i, L.origin, 12);
    if (masks8) std::memcpy(masks8 + 8 * i, L.mask, 64);
    if (values) std::memcpy(values + 512 * (size_t)g->channels * i, L.values.data(), 512 * (size_t)g->channels * 4);
  }
  return 0;
}

}  // extern "C"
