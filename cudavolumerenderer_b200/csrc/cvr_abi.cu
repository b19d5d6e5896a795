// cvr_abi.cu -- implementation of include/cvr_abi.h (libcvr_b200.so).
//
// The handle plays the role of the reference's VolPTKernelLauncher<DeviceScene>
// (RenderKernelLauncher.h:54-73) plus the device half of CudaVolPath
// (CudaVolPath.cpp): it owns the device volume (cell layouts), the path queue head,
// the counters and a stream; the caller owns the output buffer.  There is no CPU
// fallback: every entry point that needs the GPU fails with a message when CUDA does.
#include "../../include/cvr_abi.h"

#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "cvr_kernels.cuh"
#include "cvr_volume.cuh"

using namespace cvr;

namespace {

// VAR_STREAM_MK = streamingMK: Rng(c_seed + path_id) per path (StreamingVolPTmk_kernel.cuh:55),
// pull-back at scatter (:194), seed += n_paths per reset (RenderKernelLauncher.cu:480-481)
enum Variant { VAR_NAIVE = 0, VAR_REGEN = 1, VAR_STREAM = 2, VAR_STREAM_MK = 3 };

thread_local std::string g_create_error;

}  // namespace

struct cvr_renderer {
  int device = 0;
  int variant = VAR_REGEN;
  std::string kernel_name;
  std::string err;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int sm_count = 0;

  // options
  int rng_mode = RNG_XORWOW_PATH;
  int layout = LAYOUT_CELL8;        // option "layout": how cvr_set_scene lays a DENSE volume out
  int scene_layout = LAYOUT_CELL8;  // layout of the scene in place (LAYOUT_BRICK after a sparse / procedural-sparse scene)
  int exact = 0;  // 0 = fused arithmetic (queued scheduler only), 1 = the reference's operation order
  int rr = 1;
  uint32_t max_bounces = 1u << 20;
  int block = 0;  // 0 = the selected kernel's launch-bounds block (CVR_WBLOCK for sched=warp, else CVR_BLOCK)
  int blocks_per_sm = 0;  // 0 = occupancy query
  int loop_threshold = 16;
  int counters = 1;
  int sched = 3;  // 0 = lane-persistent, 1 = block-sorted wavefront, 2 = queued wavefront, 3 = warp-private wavefront
  int policy = 0; // warp scheduler batch choice (KernelParams::policy)
  int pair = -1;  // speculative second Woodcock step: -1 = default (on), 0, 1
  size_t smem_bytes = 0;  // dynamic shared memory of the selected kernel
  int warp_slots = 0;     // warp scheduler: path slots per warp (64 | 96), 0 = auto
  size_t volume_bytes = 0;  // device footprint of the density + albedo lookup layouts
  int l2_bytes = 0;
  int track_steps = 16;
  int track_min_lanes = 12;
  int tracking = 0;  // 0 = global majorant (reference), 1 = local majorant bricks
  int fix_nan = 0;
  int exit_others = 16;  // KernelParams::exit_others
  size_t l2_fetch_saved = 0;  // the device's L2 fetch granularity before option "l2_fetch" changed it (0 = untouched)
  int regen_block = 0;  // regen_order=block: path ids walk 8 x 4 pixel blocks (statistical parity; measured: no gain)
  int skip = -1;  // fetch-skip table (cvr_kernels.cuh: SkipTab): -1 = auto (on where it applies), 0, 1

  // launcher state
  KernelParams P{};
  bool scene_set = false;
  uint32_t tile_w = 0, tile_h = 0;
  uint32_t iterations = 1;
  uint32_t seed = 0;
  uint32_t sample_first = 0, sample_count = 0;
  float4* d_out = nullptr;
  uint2* trace_log = nullptr;  // cvr_trace_paths_logged: event log of the launch in flight
  uint32_t trace_log_cap = 0;

  // device memory owned by the handle
  float* d_density = nullptr;
  float* d_dcells = nullptr;
  float4* d_albedo = nullptr;
  float* d_acells = nullptr;  // AlbedoCell layout, CVR_ACELL_FLOATS floats per cell
  float* d_majorant = nullptr;
  uint32_t maj_dim[3] = {0, 0, 0};
  float* d_majorant2 = nullptr;  // second level: max over 8^3 bricks
  uint32_t maj2_dim[3] = {0, 0, 0};
  std::vector<std::pair<void*, size_t>> parked;  // freed volume blocks kept for an equal-sized request
  size_t parked_bytes = 0;
  std::unordered_map<void*, size_t> block_bytes;  // size of every live volume block
  uint8_t* d_skip = nullptr;     // fetch-skip table (global copy; the kernel stages it in shared memory)
  size_t skip_bytes = 0;         // padded to 4 bytes; 0 = the selected kernel runs without it
  uint32_t skip_shift = 0, skip_dim[3] = {0, 0, 0};
  bool skip_dirty = true;        // scene or table geometry changed since the table was built
  uint32_t* d_btable = nullptr;  // brick layout: slot table over the brick grid
  uint64_t n_bricks = 0;         // brick layout: stored bricks (without the zero brick)
  unsigned long long* d_head = nullptr;
  DeviceCounters* d_ctr = nullptr;
  bool allocated = false;

  // render_image scratch
  float4* d_tile = nullptr;
  size_t d_tile_px = 0;
  float4* d_image = nullptr;
  size_t d_image_px = 0;
  uint2* d_origins = nullptr;
  size_t d_origins_n = 0;

  // launch shape
  int grid = 0;
  int regs = 0;
  bool inited = false;

  // statistics
  uint64_t launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;
  double kernel_ms = 0.0;
};

namespace {

int fail(cvr_handle h, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h)
    h->err = buf;
  else
    g_create_error = buf;
  return 1;
}

#define CVR_CUDA(h, call)                                                               \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) return fail(h, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// CVR_TRACE_SLOW=<ms>: print host-side phases of cvr_set_scene / cvr_render_image that took longer
struct PhaseTimer {
  const char* what;
  double limit_ms;
  std::chrono::steady_clock::time_point t0;
  explicit PhaseTimer(const char* w) : what(w), limit_ms(-1.0) {
    if (const char* e = getenv("CVR_TRACE_SLOW")) limit_ms = atof(e);
    t0 = std::chrono::steady_clock::now();
  }
  void mark(const char* phase) {
    if (limit_ms < 0) return;
    auto t1 = std::chrono::steady_clock::now();
    double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    if (ms >= limit_ms) fprintf(stderr, "[cvr] %s: %s took %.2f ms\n", what, phase, ms);
    t0 = t1;
  }
};

#define CVR_CHECK_HANDLE(h) \
  if (!(h)) return fail(nullptr, "null handle")

typedef void (*kernel_fn)(const KernelParams);

template <int W>
kernel_fn pick_warp_kernel(int rng_mode, int layout, int count, int exact, int tracking, int skip) {
  if (skip) {  // fetch-skip table: fused arithmetic, global majorant, cell layouts (one 896-thread CTA per SM)
    if (exact || tracking || layout == LAYOUT_LINEAR) return nullptr;
#define CVR_KS(R, L)                                                                      \
  if (rng_mode == R && layout == L)                                                       \
    return count ? (kernel_fn)k_volpt_warp<R, L, true, true, false, W, true>              \
                 : (kernel_fn)k_volpt_warp<R, L, false, true, false, W, true>;
    CVR_KS(RNG_XORWOW_PATH, LAYOUT_CELL8)
    CVR_KS(RNG_XORWOW_PATH, LAYOUT_BRICK)
    CVR_KS(RNG_XORWOW_THREAD, LAYOUT_CELL8)
    CVR_KS(RNG_XORWOW_THREAD, LAYOUT_BRICK)
    CVR_KS(RNG_PHILOX, LAYOUT_CELL8)
    CVR_KS(RNG_PHILOX, LAYOUT_BRICK)
#undef CVR_KS
    return nullptr;
  }
  if (layout == LAYOUT_BRICK) {  // sparse bricks: fused arithmetic only, global or local majorant
    if (exact) return nullptr;   // never render exact=1 with fused arithmetic silently
#define CVR_KB(R, LOCAL)                                                                              \
  if (rng_mode == R && tracking == LOCAL)                                                             \
    return count ? (kernel_fn)k_volpt_warp<R, LAYOUT_BRICK, true, true, LOCAL != 0, W>                  \
                 : (kernel_fn)k_volpt_warp<R, LAYOUT_BRICK, false, true, LOCAL != 0, W>;
    CVR_KB(RNG_XORWOW_PATH, 0)
    CVR_KB(RNG_XORWOW_PATH, 1)
    CVR_KB(RNG_XORWOW_THREAD, 0)
    CVR_KB(RNG_XORWOW_THREAD, 1)
    CVR_KB(RNG_PHILOX, 0)
    CVR_KB(RNG_PHILOX, 1)
#undef CVR_KB
    return nullptr;
  }
  if (tracking == 1) {
    if (layout != LAYOUT_CELL8) return nullptr;
    if (rng_mode == RNG_XORWOW_PATH)
      return count ? (kernel_fn)k_volpt_warp<RNG_XORWOW_PATH, LAYOUT_CELL8, true, true, true, W>
                   : (kernel_fn)k_volpt_warp<RNG_XORWOW_PATH, LAYOUT_CELL8, false, true, true, W>;
    if (rng_mode == RNG_XORWOW_THREAD)
      return count ? (kernel_fn)k_volpt_warp<RNG_XORWOW_THREAD, LAYOUT_CELL8, true, true, true, W>
                   : (kernel_fn)k_volpt_warp<RNG_XORWOW_THREAD, LAYOUT_CELL8, false, true, true, W>;
    if (rng_mode == RNG_PHILOX)
      return count ? (kernel_fn)k_volpt_warp<RNG_PHILOX, LAYOUT_CELL8, true, true, true, W>
                   : (kernel_fn)k_volpt_warp<RNG_PHILOX, LAYOUT_CELL8, false, true, true, W>;
    return nullptr;
  }
  // counter-based stream: fused arithmetic only (its parity is statistical anyway), cell layouts
  if (rng_mode == RNG_PHILOX) {
    if (exact || layout != LAYOUT_CELL8) return nullptr;
    return count ? (kernel_fn)k_volpt_warp<RNG_PHILOX, LAYOUT_CELL8, true, true, false, W>
                 : (kernel_fn)k_volpt_warp<RNG_PHILOX, LAYOUT_CELL8, false, true, false, W>;
  }
#define CVR_K(R, L)                                                                                    \
  if (rng_mode == R && layout == L) {                                                                  \
    if (exact) return count ? (kernel_fn)k_volpt_warp<R, L, true, false, false, W> : (kernel_fn)k_volpt_warp<R, L, false, false, false, W>; \
    return count ? (kernel_fn)k_volpt_warp<R, L, true, true, false, W> : (kernel_fn)k_volpt_warp<R, L, false, true, false, W>;           \
  }
  CVR_K(RNG_XORWOW_PATH, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_PATH, LAYOUT_LINEAR)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_LINEAR)
#undef CVR_K
  return nullptr;
}

// instantiations with the per-path event log (cvr_trace_paths_logged): warp scheduler, per-path
// XORWOW streams, global majorant, no skip table, 64 slots
kernel_fn pick_logged_kernel(int layout, int exact) {
  if (layout == LAYOUT_CELL8)
    return exact ? (kernel_fn)k_volpt_warp<RNG_XORWOW_PATH, LAYOUT_CELL8, true, false, false, 64, false, true>
                 : (kernel_fn)k_volpt_warp<RNG_XORWOW_PATH, LAYOUT_CELL8, true, true, false, 64, false, true>;
  if (layout == LAYOUT_BRICK && !exact)
    return (kernel_fn)k_volpt_warp<RNG_XORWOW_PATH, LAYOUT_BRICK, true, true, false, 64, false, true>;
  return nullptr;
}

kernel_fn pick_kernel(int sched, int rng_mode, int layout, int count, int exact = 1, int tracking = 0, int wslots = 64,
                      int skip = 0) {
  if (sched == 3)
    return wslots == 96 ? pick_warp_kernel<96>(rng_mode, layout, count, exact, tracking, skip)
                        : pick_warp_kernel<64>(rng_mode, layout, count, exact, tracking, skip);
  if (tracking == 1) {
    if (sched != 2 || layout != LAYOUT_CELL8) return nullptr;
    if (rng_mode == RNG_XORWOW_PATH)
      return count ? (kernel_fn)k_volpt_queued<RNG_XORWOW_PATH, LAYOUT_CELL8, true, true, true>
                   : (kernel_fn)k_volpt_queued<RNG_XORWOW_PATH, LAYOUT_CELL8, false, true, true>;
    if (rng_mode == RNG_XORWOW_THREAD)
      return count ? (kernel_fn)k_volpt_queued<RNG_XORWOW_THREAD, LAYOUT_CELL8, true, true, true>
                   : (kernel_fn)k_volpt_queued<RNG_XORWOW_THREAD, LAYOUT_CELL8, false, true, true>;
    return nullptr;
  }
#define CVR_K(R, L)                                      \
  if (sched == 0 && rng_mode == R && layout == L)        \
    return count ? (kernel_fn)k_volpt<R, L, true> : (kernel_fn)k_volpt<R, L, false>;
  CVR_K(RNG_XORWOW_PATH, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_PATH, LAYOUT_LINEAR)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_LINEAR)
  CVR_K(RNG_PHILOX, LAYOUT_CELL8)
  CVR_K(RNG_PHILOX, LAYOUT_LINEAR)
#undef CVR_K
#define CVR_K(R, L)                                      \
  if (sched == 1 && rng_mode == R && layout == L)        \
    return count ? (kernel_fn)k_volpt_sorted<R, L, true> : (kernel_fn)k_volpt_sorted<R, L, false>;
  CVR_K(RNG_XORWOW_PATH, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_PATH, LAYOUT_LINEAR)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_LINEAR)
#undef CVR_K
#define CVR_K(R, L)                                                                                    \
  if (sched == 2 && rng_mode == R && layout == L) {                                                    \
    if (exact) return count ? (kernel_fn)k_volpt_queued<R, L, true, false> : (kernel_fn)k_volpt_queued<R, L, false, false>; \
    return count ? (kernel_fn)k_volpt_queued<R, L, true, true> : (kernel_fn)k_volpt_queued<R, L, false, true>;           \
  }
  CVR_K(RNG_XORWOW_PATH, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_PATH, LAYOUT_LINEAR)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_CELL8)
  CVR_K(RNG_XORWOW_THREAD, LAYOUT_LINEAR)
#undef CVR_K
  return nullptr;
}

int set_device(cvr_handle h) {
  CVR_CUDA(h, cudaSetDevice(h->device));
  return 0;
}

// Volume memory comes from the device's stream-ordered pool with an unlimited release
// threshold (cudaMalloc / cudaFree of 10^8-byte blocks cost anything between 5 and 250 ms per
// cvr_set_scene on this driver), and blocks a scene change frees are first parked in the handle
// and handed back to a request of exactly the same size: a re-upload of a same-shaped scene --
// an animated volume, the bench's end-to-end step -- then makes NO allocator call at all.
// (Even cudaMallocAsync stalls for 200-450 ms when an nvidia-smi query runs beside it: measured.)
// Everything is ordered on the handle's stream, so a parked block can be rewritten at once.
cudaError_t vol_alloc(cvr_handle h, void** p, size_t bytes) {
  bytes = bytes ? bytes : 1;
  for (size_t i = 0; i < h->parked.size(); ++i)
    if (h->parked[i].second == bytes) {
      *p = h->parked[i].first;
      h->parked_bytes -= bytes;
      h->parked.erase(h->parked.begin() + (long)i);
      return cudaSuccess;
    }
  cudaError_t e = cudaMallocAsync(p, bytes, h->stream);
  if (e == cudaSuccess) h->block_bytes[*p] = bytes;
  return e;
}
void vol_free(cvr_handle h, void* p) {
  if (!p) return;
  auto it = h->block_bytes.find(p);
  const size_t bytes = it == h->block_bytes.end() ? 0 : it->second;
  if (bytes && bytes <= (512ull << 20) && h->parked_bytes + bytes <= (1ull << 30) && h->parked.size() < 16) {
    h->parked.emplace_back(p, bytes);
    h->parked_bytes += bytes;
    return;
  }
  if (it != h->block_bytes.end()) h->block_bytes.erase(it);
  cudaFreeAsync(p, h->stream);
}
void vol_flush_parked(cvr_handle h) {
  for (auto& b : h->parked) {
    h->block_bytes.erase(b.first);
    cudaFreeAsync(b.first, h->stream);
  }
  h->parked.clear();
  h->parked_bytes = 0;
}

void free_volume(cvr_handle h) {
  vol_free(h, h->d_density);
  vol_free(h, h->d_dcells);
  vol_free(h, h->d_albedo);
  vol_free(h, h->d_acells);
  vol_free(h, h->d_majorant);
  vol_free(h, h->d_majorant2);
  h->d_majorant2 = nullptr;
  vol_free(h, h->d_btable);
  vol_free(h, h->d_skip);
  h->d_skip = nullptr;
  h->skip_dirty = true;
  h->d_majorant = nullptr;
  h->d_btable = nullptr;
  h->n_bricks = 0;
  h->d_density = h->d_dcells = nullptr;
  h->d_albedo = nullptr;
  h->d_acells = nullptr;
  h->scene_set = false;
  // a multi-GB volume must not stay parked in the pool (other allocators of the process --
  // torch, the caller's cudaMalloc -- cannot see it): give everything above 1 GiB back
  cudaMemPool_t pool;
  unsigned long long reserved = 0;
  if (cudaDeviceGetDefaultMemPool(&pool, h->device) == cudaSuccess &&
      cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
      reserved > (2ull << 30)) {
    cudaStreamSynchronize(h->stream);
    cudaMemPoolTrimTo(pool, 2ull << 30);
  }
}

int ensure_allocated(cvr_handle h) {
  if (h->allocated) return 0;
  // stream-ordered pool (like the volumes): a renderer built per trial (cvr_render, the reference's runTest) then
  // costs no cudaMalloc -- five of them per handle, serialised across the 8 host threads of a device group, were
  // 17 ms of a 29 ms C3 render on 8 GPUs
  CVR_CUDA(h, cudaMallocAsync((void**)&h->d_head, sizeof(unsigned long long), h->stream));
  CVR_CUDA(h, cudaMallocAsync((void**)&h->d_ctr, sizeof(DeviceCounters), h->stream));
  CVR_CUDA(h, cudaMemsetAsync(h->d_head, 0, sizeof(unsigned long long), h->stream));
  CVR_CUDA(h, cudaMemsetAsync(h->d_ctr, 0, sizeof(DeviceCounters), h->stream));
  h->allocated = true;
  return 0;
}

// slots per warp of the warp-private scheduler: "auto" = 96 while the density lookup layout
// (the stream with ~90 % of the lookups) fits the L2 (fuller batches win), 64 beyond (the L1
// left next to the slots wins); measurements in cvr_kernels.cuh
int effective_wslots(cvr_handle h) {
  if (h->warp_slots) return h->warp_slots;
  return (h->volume_bytes && h->volume_bytes <= (size_t)h->l2_bytes) ? 96 : 64;
}

// The speculative second step doubles the loads in flight per lane (latency) at the price of
// ~14 % extra lookups (bandwidth).  Measured with the warp scheduler (1024^2 x 64 spp, pair
// off / on, Msamples/s): hetvol 889 / 973, manix 1931 / 2290, fbm 512^3 1029 / 1116, fbm 1024^3
// 1026 / 1050, sparse 2048^3 1612 / 2058 -- on by default, also for HBM-resident volumes.
int effective_pair(cvr_handle h) { return h->pair >= 0 ? h->pair : 1; }

// The fetch-skip table applies to the fused global-majorant loop of the warp scheduler over a
// cell layout; "auto" turns it on there.
bool skip_wanted(cvr_handle h) {
  if (h->sched != 3 || h->exact || h->tracking || h->scene_layout == LAYOUT_LINEAR) return false;
  if (!h->d_majorant || !h->maj_dim[0]) return false;
  // auto: on for volumes beyond the L2.  Measured on B200 (1024^2 x 16 spp, Msamples/s off / on):
  // manix 2526 / 2744, fbm 512^3 1079 / 1480, sparse 1024^3 2447 / 2475 -- but bucky 5679 / 5196 and
  // hetvol 985 / 945: an L2-resident volume has no bandwidth to save and pays the ~30 extra
  // instructions per pair of steps.
  if (h->skip < 0) return h->volume_bytes > (size_t)h->l2_bytes;
  return h->skip != 0;
}

int effective_block(cvr_handle h) {
  if (h->sched == 3 && h->skip_bytes) return CVR_WSKIP_BLOCK;  // the table offset is compiled for this CTA size
  const int cap = h->sched == 3 ? CVR_WBLOCK : CVR_BLOCK;
  return (h->block > 0 && h->block <= cap) ? h->block : cap;
}

// Table geometry: the finest brick edge (8 << e cells) whose byte-per-brick table fits the shared
// memory left beside the slots of ONE CTA of CVR_WSKIP_BLOCK threads; skip_bytes = 0 when no
// level up to 64^3-cell bricks fits (the kernel then runs without the table).
void plan_skip_table(cvr_handle h, size_t smem_optin) {
  h->skip_bytes = 0;
  if (!skip_wanted(h)) return;
  const size_t slots = warp_sched_smem_bytes(CVR_WSKIP_BLOCK, effective_wslots(h), 0, warp_slot_bytes(h->rng_mode));
  if (slots + 64 > smem_optin) return;
  const size_t budget = smem_optin - slots;
  for (uint32_t e = 0; e <= 3; ++e) {
    uint32_t d[3];
    for (int a = 0; a < 3; ++a) d[a] = (h->maj_dim[a] + (1u << e) - 1) >> e;
    const size_t n = ((size_t)d[0] * d[1] * d[2] + 3) & ~(size_t)3;
    if (n <= budget) {
      if (h->skip_shift != CVR_BRICK_LOG2 + e || h->skip_dim[0] != d[0] || h->skip_dim[1] != d[1] || h->skip_dim[2] != d[2])
        h->skip_dirty = true;
      h->skip_shift = CVR_BRICK_LOG2 + e;
      h->skip_dim[0] = d[0], h->skip_dim[1] = d[1], h->skip_dim[2] = d[2];
      h->skip_bytes = n;
      return;
    }
  }
}

// (re)build the table for the scene in place; needs P.inv.sig_ratio (fill_track_inv)
int build_skip_table(cvr_handle h) {
  if (!h->skip_bytes || !h->skip_dirty) return 0;
  vol_free(h, h->d_skip);
  h->d_skip = nullptr;
  CVR_CUDA(h, vol_alloc(h, (void**)&h->d_skip, h->skip_bytes));
  CVR_CUDA(h, cudaMemsetAsync(h->d_skip, 0xff, h->skip_bytes, h->stream));  // padding bytes: never skip
  const uint32_t n = h->skip_dim[0] * h->skip_dim[1] * h->skip_dim[2];
  k_build_skip_table<<<(n + 127) / 128, 128, 0, h->stream>>>(h->d_majorant, h->maj_dim[0], h->maj_dim[1], h->maj_dim[2],
                                                             h->skip_shift - CVR_BRICK_LOG2, h->skip_dim[0], h->skip_dim[1],
                                                             h->skip_dim[2], h->P.inv.sig_ratio, h->d_skip);
  CVR_CUDA(h, cudaGetLastError());
  h->skip_dirty = false;
  return 0;
}

int ensure_init(cvr_handle h) {
  if (h->inited) return 0;
  {
    int optin = 0;
    CVR_CUDA(h, cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    plan_skip_table(h, (size_t)optin);
  }
  kernel_fn k = pick_kernel(h->sched, h->rng_mode, h->scene_layout, h->counters, h->exact, h->tracking, effective_wslots(h),
                            h->skip_bytes != 0);
  if (!k)
    return fail(h, "no kernel for sched=%d rng=%d layout=%d exact=%d tracking=%d (rng=philox needs sched=warp with exact=0 and "
                "layout=cell8, or sched=lane; sparse scenes need sched=warp, exact=0; tracking=local needs sched=warp|queued)",
                h->sched, h->rng_mode, h->scene_layout, h->exact, h->tracking);
  cudaFuncAttributes fa;
  CVR_CUDA(h, cudaFuncGetAttributes(&fa, (const void*)k));
  h->regs = fa.numRegs;
  // the warp-private scheduler keeps its path slots in DYNAMIC shared memory (may exceed 48 KB)
  h->smem_bytes = h->sched == 3 ? warp_sched_smem_bytes(effective_block(h), effective_wslots(h), h->skip_bytes,
                                                        warp_slot_bytes(h->rng_mode))
                                : 0;
  if (h->smem_bytes)
    CVR_CUDA(h, cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
  int per_sm = 0;
  CVR_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, effective_block(h), h->smem_bytes));
  if (per_sm < 1) return fail(h, "kernel does not fit an SM at block size %d", effective_block(h));
  if (h->blocks_per_sm > 0 && h->blocks_per_sm < per_sm) per_sm = h->blocks_per_sm;
  h->grid = per_sm * h->sm_count;
  h->inited = true;
  return 0;
}

void collect_timing(cvr_handle h) {
  for (auto& ev : h->timing) {
    float ms = 0.f;
    if (cudaEventSynchronize(ev.second) == cudaSuccess &&
        cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess)
      h->kernel_ms += ms;
    cudaEventDestroy(ev.first);
    cudaEventDestroy(ev.second);
  }
  h->timing.clear();
}

// Woodcock-loop invariants, same fp32 operations the kernels used to perform per thread
void fill_track_inv(KernelParams& P) {
  const MediumParams& m = P.med;
  TrackInv& I = P.inv;
  volatile float prod = m.scale * m.max_density;  // keep the two roundings separate
  I.inv_max_sigmat = 1.0f / prod;
  volatile float ex = m.box_max.x - m.box_min.x, ey = m.box_max.y - m.box_min.y, ez = m.box_max.z - m.box_min.z;
  I.qx = m.box_min.x / ex, I.qy = m.box_min.y / ey, I.qz = m.box_min.z / ez;
  I.rx = (float)(uint32_t)(m.dnx - 1), I.ry = (float)(uint32_t)(m.dny - 1), I.rz = (float)(uint32_t)(m.dnz - 1);
  I.nx = m.dnx, I.ny = m.dny, I.nz = m.dnz;
  I.sy = m.dnx + 1, I.sz = (uint32_t)(m.dnx + 1) * (uint32_t)(m.dny + 1);
  I.nqrx = -I.qx * I.rx, I.nqry = -I.qy * I.ry, I.nqrz = -I.qz * I.rz;
  I.sig_ratio = m.scale * I.inv_max_sigmat;
  I.aix = 1.0f / ex, I.aiy = 1.0f / ey, I.aiz = 1.0f / ez;
  I.neg_ln2_inv_sigmat = -0.69314718055994530942f * I.inv_max_sigmat;
}

void fill_majorant(cvr_handle h) {
  h->P.inv.majorant = h->d_majorant;
  h->P.inv.mx = h->maj_dim[0], h->P.inv.my = h->maj_dim[1], h->P.inv.mz = h->maj_dim[2];
  h->P.inv.majorant2 = h->d_majorant2;
  h->P.inv.m2x = h->maj2_dim[0], h->P.inv.m2y = h->maj2_dim[1], h->P.inv.m2z = h->maj2_dim[2];
}

// second majorant level from the brick majorants (h->d_majorant, h->maj_dim)
int build_majorant2(cvr_handle h) {
  for (int a = 0; a < 3; ++a) h->maj2_dim[a] = (h->maj_dim[a] + 7) / 8;
  const size_t n2 = (size_t)h->maj2_dim[0] * h->maj2_dim[1] * h->maj2_dim[2];
  CVR_CUDA(h, vol_alloc(h, (void**)&h->d_majorant2, n2 * sizeof(float)));
  k_build_majorant2<<<(unsigned)((n2 + 127) / 128), 128, 0, h->stream>>>(h->d_majorant, h->maj_dim[0], h->maj_dim[1],
                                                                        h->maj_dim[2], h->d_majorant2, h->maj2_dim[0],
                                                                        h->maj2_dim[1], h->maj2_dim[2]);
  CVR_CUDA(h, cudaGetLastError());
  return 0;
}

// fills the per-launch part of the kernel parameters and launches
int launch(cvr_handle h, float4* out, uint32_t out_stride, int out_full, const uint2* origins,
           uint32_t n_launch_tiles, uint32_t tile_first, uint32_t tile_stride, uint32_t seed,
           uint32_t seed_step, float4* per_path, unsigned long long path_begin,
           unsigned long long path_end) {
  if (!h->scene_set) return fail(h, "launch before cvr_set_scene");
  if (h->tile_w == 0 || h->tile_h == 0) return fail(h, "launch before cvr_set_resolution");
  if (path_end < path_begin) return fail(h, "launch: empty or inverted path range [%llu, %llu)", path_begin, path_end);
  if (set_device(h)) return 1;
  if (ensure_allocated(h) || ensure_init(h)) return 1;
  KernelParams& P = h->P;
  fill_track_inv(P);
  fill_majorant(h);
  P.npix = (uint32_t)(P.cam.res_x * P.cam.res_y);  // (uint)(c_resolution.x * c_resolution.y)
  P.tile_w = (uint32_t)P.cam.res_x;
  P.path_begin = path_begin;
  P.path_end = path_end;
  P.tile_origins = origins;
  P.n_launch_tiles = n_launch_tiles;
  P.tile_first = tile_first;
  P.tile_stride = tile_stride;
  P.seed = seed;
  P.seed_step = seed_step;
  P.out = out;
  P.out_stride = out_stride;
  P.out_full = out_full;
  P.per_path = per_path;
  P.path_log = per_path ? h->trace_log : nullptr;  // the event log belongs to cvr_trace_paths_logged only
  P.log_cap = h->trace_log_cap;
  P.head = h->d_head;
  P.ctr = h->d_ctr;
  P.max_bounces = h->max_bounces;
  P.loop_threshold = h->loop_threshold;
  P.track_steps = h->track_steps;
  P.track_min_lanes = h->track_min_lanes;
  P.fix_nan = h->fix_nan;
  P.policy = h->policy;
  P.exit_others = h->exit_others;
  // the block order needs whole 8 x 4 blocks; any other tile shape keeps the row order
  P.regen_block = (h->regen_block && h->tile_w % 8u == 0 && h->tile_h % 4u == 0) ? 1u : 0u;
  P.pair = effective_pair(h);
  P.rr = h->rr;
  P.pullback = (h->variant != VAR_REGEN) ? 1 : 0;
  P.rr_after_escape = (h->variant != VAR_NAIVE && h->rng_mode == RNG_XORWOW_THREAD) ? 1 : 0;
  CVR_CUDA(h, cudaMemsetAsync(h->d_head, 0, sizeof(unsigned long long), h->stream));
  if (build_skip_table(h)) return 1;
  P.skip_tab = h->d_skip;
  P.skip_n = (uint32_t)h->skip_bytes;
  P.skip_shift = h->skip_shift;
  P.skip_bx = h->skip_dim[0];
  P.skip_bxy = h->skip_dim[0] * h->skip_dim[1];
  kernel_fn k = pick_kernel(h->sched, h->rng_mode, h->scene_layout, h->counters, h->exact, h->tracking, effective_wslots(h),
                            h->skip_bytes != 0);
  int grid = h->grid, block = effective_block(h);
  size_t smem = h->smem_bytes;
  if (P.path_log) {
    // cvr_trace_paths_logged: the instantiation that keeps the event log (same estimator blocks,
    // same arithmetic mode; 64 slots per warp, no skip table -- neither changes any path)
    k = pick_logged_kernel(h->scene_layout, h->exact);
    if (!k) return fail(h, "cvr_trace_paths_logged: no logging kernel for layout=%d exact=%d", h->scene_layout, h->exact);
    block = CVR_WBLOCK;
    smem = warp_sched_smem_bytes(block, 64, 0);
    P.skip_tab = nullptr, P.skip_n = 0;
    CVR_CUDA(h, cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CVR_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)k, block, smem));
    if (per_sm < 1) return fail(h, "cvr_trace_paths_logged: the logging kernel does not fit an SM");
    grid = per_sm * h->sm_count;
  }
  cudaEvent_t e0, e1;
  CVR_CUDA(h, cudaEventCreate(&e0));
  CVR_CUDA(h, cudaEventCreate(&e1));
  CVR_CUDA(h, cudaEventRecord(e0, h->stream));
  k<<<grid, block, smem, h->stream>>>(P);
  CVR_CUDA(h, cudaGetLastError());
  CVR_CUDA(h, cudaEventRecord(e1, h->stream));
  h->timing.emplace_back(e0, e1);
  if (h->timing.size() > 4096) collect_timing(h);
  h->launches++;
  return 0;
}

// a sample range outside the iteration count is a caller bug, not something to clamp silently
int check_sample_range(cvr_handle h) {
  if (h->sample_first == 0 && h->sample_count == 0) return 0;
  if (h->sample_first >= h->iterations)
    return fail(h, "sample range: first sample %u is not below the iteration count %u", h->sample_first, h->iterations);
  if ((unsigned long long)h->sample_first + h->sample_count > (unsigned long long)h->iterations)
    return fail(h, "sample range: samples [%u, %llu) exceed the iteration count %u", h->sample_first,
                (unsigned long long)h->sample_first + h->sample_count, h->iterations);
  return 0;
}

void path_range(cvr_handle h, unsigned long long& b, unsigned long long& e) {
  // 64-bit throughout: first is clamped to the iteration count, the count to what is left of it
  // (a bad range must never give path_end < path_begin: the persistent kernels' `per_tile` is unsigned)
  const unsigned long long npix = (unsigned long long)(uint32_t)(h->P.cam.res_x * h->P.cam.res_y);
  const unsigned long long iters = h->iterations;
  const unsigned long long first = std::min<unsigned long long>(h->sample_first, iters);
  const unsigned long long left = iters - first;
  const unsigned long long count = h->sample_count ? std::min<unsigned long long>(h->sample_count, left) : left;
  b = npix * first;
  e = npix * (first + count);
}

}  // namespace

// =========================================================================== ABI
extern "C" {

int cvr_abi_version(void) { return CVR_ABI_VERSION; }

const char* cvr_last_error(cvr_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int cvr_create(const char* kernel_name, int device, cvr_handle* out) {
  if (!out) return fail(nullptr, "cvr_create: out is null");
  *out = nullptr;
  if (!kernel_name) return fail(nullptr, "cvr_create: kernel name is null");
  int variant;
  std::string k(kernel_name);
  if (k == "naiveSK" || k == "naive")
    variant = VAR_NAIVE;
  else if (k == "regenerationSK")
    variant = VAR_REGEN;
  else if (k == "streamingSK")
    variant = VAR_STREAM;
  else if (k == "sortingSK")  // same estimator, streams and seed rule as streamingSK (SortingVolPTsk_kernel.cuh:227-230,
    variant = VAR_STREAM;     // :314; RenderKernelLauncher.cu:664-665); its Morton ordering is scheduling, which is ours
  else if (k == "streamingMK")
    variant = VAR_STREAM_MK;
  else if (k == "naiveMK")
    return fail(nullptr,
                "cvr_create: kernel 'naiveMK' is not offered: its per-bounce reseeding and first-hit handling "
                "(NaiveVolPTmk_kernel.cuh:32-75,90) make it a different estimator variant; use naiveSK");
  else
    return fail(nullptr,
                "cvr_create: unknown kernel '%s' (naiveSK | regenerationSK | streamingSK | streamingMK | sortingSK)",
                kernel_name);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, "cvr_create: no CUDA device (%s); there is no CPU fallback",
                cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail(nullptr, "cvr_create: device %d of %d", device, n_dev);
  cvr_handle h = new cvr_renderer();
  h->device = device;
  h->variant = variant;
  h->kernel_name = k;
  if (cudaSetDevice(device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    fail(nullptr, "cvr_create: cannot create a stream on device %d: %s", device,
         cudaGetErrorString(cudaGetLastError()));
    delete h;
    return 1;
  }
  h->stream = h->own_stream;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  h->sm_count = prop.multiProcessorCount;
  h->l2_bytes = prop.l2CacheSize;
  {
    cudaMemPool_t pool;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  if (prop.major < 10) {
    fail(nullptr, "cvr_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
         prop.major, prop.minor);
    cudaStreamDestroy(h->own_stream);
    delete h;
    return 1;
  }
  // reference defaults
  h->P.med.hg_g = 0.f;
  h->P.med.alpha_x = h->P.med.alpha_y = 0.1f;
  h->P.med.eta = 1.05f / 1.01f;
  // naiveSK: Rng(tid) with no seed (Q7); streams are per path by construction
  *out = h;
  return 0;
}

int cvr_release(cvr_handle h) {
  CVR_CHECK_HANDLE(h);
  if (set_device(h)) return 1;
  cudaStreamSynchronize(h->stream);
  collect_timing(h);
  free_volume(h);
  vol_flush_parked(h);
  cudaStreamSynchronize(h->stream);
  for (void* p : {(void*)h->d_head, (void*)h->d_ctr, (void*)h->d_tile, (void*)h->d_image, (void*)h->d_origins})
    if (p) cudaFreeAsync(p, h->stream);
  cudaStreamSynchronize(h->stream);
  h->d_head = nullptr, h->d_ctr = nullptr, h->d_tile = nullptr, h->d_image = nullptr;
  h->d_origins = nullptr;
  h->d_tile_px = h->d_image_px = h->d_origins_n = 0;
  h->allocated = false;
  return 0;
}

int cvr_destroy(cvr_handle h) {
  if (!h) return 0;
  cvr_release(h);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return 0;
}

int cvr_set_option(cvr_handle h, const char* key, const char* value) {
  CVR_CHECK_HANDLE(h);
  if (!key || !value) return fail(h, "cvr_set_option: null key/value");
  std::string k(key), v(value);
  if (k == "rng") {
    if (v == "xorwow-path" || v == "xorwow")
      h->rng_mode = RNG_XORWOW_PATH;
    else if (v == "xorwow-thread")
      h->rng_mode = RNG_XORWOW_THREAD;
    else if (v == "philox")
      h->rng_mode = RNG_PHILOX;
    else
      return fail(h, "rng: unknown value '%s'", value);
    if (h->variant == VAR_NAIVE && h->rng_mode == RNG_XORWOW_THREAD)
      return fail(h, "rng=xorwow-thread is not a naiveSK mode (NaiveVolPTsk_kernel.cuh:22 seeds per path)");
    h->inited = false;
  } else if (k == "layout") {
    if (v == "cell8")
      h->layout = LAYOUT_CELL8;
    else if (v == "linear")
      h->layout = LAYOUT_LINEAR;
    else
      return fail(h, "layout: unknown value '%s'", value);
    if (h->scene_set) return fail(h, "layout must be chosen before cvr_set_scene");
    h->inited = false;
  } else if (k == "tracking") {
    if (v == "global")
      h->tracking = 0;
    else if (v == "local")
      h->tracking = 1;
    else
      return fail(h, "tracking: unknown value '%s' (global | local)", value);
    h->inited = false;
  } else if (k == "exact") {
    h->exact = atoi(value) ? 1 : 0;
    h->inited = false;
  } else if (k == "russian_roulette") {
    h->rr = atoi(value) ? 1 : 0;
  } else if (k == "max_bounces") {
    h->max_bounces = (uint32_t)strtoul(value, nullptr, 10);
  } else if (k == "block") {
    int b = atoi(value);
    if (b < 32 || b > CVR_BLOCK || (b % 32)) return fail(h, "block must be a multiple of 32 in [32,%d]", CVR_BLOCK);
    h->block = b;
    h->inited = false;
  } else if (k == "blocks_per_sm") {
    h->blocks_per_sm = atoi(value);
    h->inited = false;
  } else if (k == "loop_threshold") {
    int t = atoi(value);
    if (t < 1) return fail(h, "loop_threshold must be >= 1");
    h->loop_threshold = t;
  } else if (k == "sched") {
    if (v == "lane")
      h->sched = 0;
    else if (v == "sorted")
      h->sched = 1;
    else if (v == "queued")
      h->sched = 2;
    else if (v == "warp")
      h->sched = 3;
    else
      return fail(h, "sched: unknown value '%s' (lane | sorted | queued | warp)", value);
    h->inited = false;
  } else if (k == "policy") {
    h->policy = atoi(value);
  } else if (k == "pair") {
    h->pair = v == "auto" ? -1 : (atoi(value) ? 1 : 0);
  } else if (k == "exit_others") {
    h->exit_others = atoi(value);
  } else if (k == "regen_order") {
    if (v != "row" && v != "block") return fail(h, "regen_order: unknown value '%s' (row | block)", value);
    h->regen_block = v == "block" ? 1 : 0;
  } else if (k == "skip") {
    h->skip = v == "auto" ? -1 : (atoi(value) ? 1 : 0);
    h->inited = false;
  } else if (k == "l2_fetch") {
    // cudaLimitMaxL2FetchGranularity (a device-wide HINT): bytes the L2 fetches from DRAM per missing sector.
    // The lookup gathers ONE 32-byte sector per step at an unpredictable address; ncu on fBm 1024^3 showed
    // 3.4 DRAM sectors read per sector the kernel asked for at the driver's default.
    if (set_device(h)) return 1;
    if (v == "default") {
      if (h->l2_fetch_saved) CVR_CUDA(h, cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, h->l2_fetch_saved));
    } else {
      const int g = atoi(value);
      if (g != 32 && g != 64 && g != 128) return fail(h, "l2_fetch: unknown value '%s' (32 | 64 | 128 | default)", value);
      if (!h->l2_fetch_saved) {
        size_t cur = 0;
        CVR_CUDA(h, cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity));
        h->l2_fetch_saved = cur ? cur : 64;
      }
      CVR_CUDA(h, cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g));
    }
  } else if (k == "warp_slots") {
    if (v == "auto")
      h->warp_slots = 0;
    else if (v == "64" || v == "96")
      h->warp_slots = atoi(value);
    else
      return fail(h, "warp_slots: unknown value '%s' (auto | 64 | 96)", value);
    h->inited = false;
  } else if (k == "track_steps") {
    int t = atoi(value);
    if (t < 1) return fail(h, "track_steps must be >= 1");
    h->track_steps = t;
  } else if (k == "track_min_lanes") {
    h->track_min_lanes = atoi(value);
  } else if (k == "fix_nan") {
    h->fix_nan = atoi(value) ? 1 : 0;
  } else if (k == "counters") {
    h->counters = atoi(value) ? 1 : 0;
    h->inited = false;
  } else {
    return fail(h, "unknown option '%s'", key);
  }
  return 0;
}

int cvr_get_option(cvr_handle h, const char* key, char* value, size_t cap) {
  CVR_CHECK_HANDLE(h);
  if (!key || !value || cap == 0) return fail(h, "cvr_get_option: bad arguments");
  std::string k(key), v;
  if (k == "rng")
    v = h->rng_mode == RNG_XORWOW_PATH ? "xorwow-path" : h->rng_mode == RNG_XORWOW_THREAD ? "xorwow-thread" : "philox";
  else if (k == "layout")
    v = h->layout == LAYOUT_CELL8 ? "cell8" : "linear";
  else if (k == "tracking")
    v = h->tracking ? "local" : "global";
  else if (k == "exact")
    v = std::to_string(h->sched >= 2 ? h->exact : 1);
  else if (k == "russian_roulette")
    v = std::to_string(h->rr);
  else if (k == "max_bounces")
    v = std::to_string(h->max_bounces);
  else if (k == "block")
    v = std::to_string(effective_block(h));
  else if (k == "blocks_per_sm")
    v = std::to_string(h->blocks_per_sm);
  else if (k == "loop_threshold")
    v = std::to_string(h->loop_threshold);
  else if (k == "counters")
    v = std::to_string(h->counters);
  else if (k == "sched")
    v = h->sched == 3 ? "warp" : h->sched == 2 ? "queued" : h->sched ? "sorted" : "lane";
  else if (k == "policy")
    v = std::to_string(h->policy);
  else if (k == "pair")
    v = std::to_string(effective_pair(h));
  else if (k == "exit_others")
    v = std::to_string(h->exit_others);
  else if (k == "regen_order")
    v = h->regen_block ? "block" : "row";
  else if (k == "skip")  // "0" or the brick edge in cells the table was planned with (after the first launch / init)
    v = h->skip_bytes ? std::to_string(1u << h->skip_shift) : std::string(skip_wanted(h) && !h->inited ? "auto" : "0");
  else if (k == "warp_slots")
    v = std::to_string(effective_wslots(h));
  else if (k == "l2_fetch") {
    size_t cur = 0;
    if (set_device(h)) return 1;
    CVR_CUDA(h, cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity));
    v = std::to_string(cur);
  }
  else if (k == "track_steps")
    v = std::to_string(h->track_steps);
  else if (k == "track_min_lanes")
    v = std::to_string(h->track_min_lanes);
  else if (k == "fix_nan")
    v = std::to_string(h->fix_nan);
  else if (k == "kernel")
    v = h->kernel_name;
  else
    return fail(h, "unknown option '%s'", key);
  snprintf(value, cap, "%s", v.c_str());
  return 0;
}

int cvr_set_stream(cvr_handle h, void* cuda_stream) {
  CVR_CHECK_HANDLE(h);
  cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  if (next != h->stream) {
    // the handle's buffers come from the stream-ordered pool: what was allocated (or freed) in the order of the old
    // stream must be settled before the new one touches it
    if (set_device(h)) return 1;
    cudaStreamSynchronize(h->stream);
  }
  h->stream = next;
  return 0;
}

int cvr_set_scene(cvr_handle h, const cvr_scene_desc* s) {
  CVR_CHECK_HANDLE(h);
  if (!s || !s->density) return fail(h, "cvr_set_scene: null scene/density");
  for (int i = 0; i < 3; ++i) {
    if (s->density_dim[i] < 2) return fail(h, "cvr_set_scene: density dims must be >= 2");
    if (s->albedo && s->albedo_dim[i] < 2) return fail(h, "cvr_set_scene: albedo dims must be >= 2");
    if (!(s->box_max[i] > s->box_min[i])) return fail(h, "cvr_set_scene: empty box");
  }
  if (!(s->scale > 0.f) || !(s->max_density > 0.f))
    return fail(h, "cvr_set_scene: scale and max_density must be positive");
  if (set_device(h)) return 1;
  PhaseTimer pt("cvr_set_scene");
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  pt.mark("sync");
  free_volume(h);
  pt.mark("free_volume");
  h->scene_layout = h->layout;
  MediumParams& m = h->P.med;
  m.btable = nullptr, m.bmx = m.bmy = m.bmz = 0;
  m.box_min = V3{s->box_min[0], s->box_min[1], s->box_min[2]};
  m.box_max = V3{s->box_max[0], s->box_max[1], s->box_max[2]};
  m.scale = s->scale;
  m.max_density = s->max_density;
  m.hg_g = s->hg_g;
  m.alpha_x = s->ggx_alpha[0] > 0.f ? s->ggx_alpha[0] : 0.1f;
  m.alpha_y = s->ggx_alpha[1] > 0.f ? s->ggx_alpha[1] : 0.1f;
  m.eta = s->ggx_eta > 0.f ? s->ggx_eta : 1.05f / 1.01f;
  m.dnx = s->density_dim[0], m.dny = s->density_dim[1], m.dnz = s->density_dim[2];
  const cudaMemcpyKind kind = s->density_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  size_t nd = (size_t)m.dnx * m.dny * m.dnz;
  CVR_CUDA(h, vol_alloc(h, (void**)&h->d_density, nd * sizeof(float)));
  CVR_CUDA(h, cudaMemcpyAsync(h->d_density, s->density, nd * sizeof(float), kind, h->stream));
  const int bt = 256;
  if (h->layout == LAYOUT_CELL8) {
    size_t ncell = (size_t)(m.dnx + 1) * (m.dny + 1) * (m.dnz + 1);
    if (ncell >= (1ull << 32)) return fail(h, "cvr_set_scene: density grid too large for 32-bit cell indices");
    CVR_CUDA(h, vol_alloc(h, (void**)&h->d_dcells, ncell * 8 * sizeof(float)));
    int g = (int)std::min<size_t>((ncell + bt - 1) / bt, (size_t)h->sm_count * 32);
    k_build_density_cells<<<g, bt, 0, h->stream>>>(h->d_density, m.dnx, m.dny, m.dnz, (float4*)h->d_dcells);
    CVR_CUDA(h, cudaGetLastError());
    // majorant bricks for tracking=local
    h->maj_dim[0] = (uint32_t)(m.dnx + 1 + CVR_BRICK - 1) / CVR_BRICK;
    h->maj_dim[1] = (uint32_t)(m.dny + 1 + CVR_BRICK - 1) / CVR_BRICK;
    h->maj_dim[2] = (uint32_t)(m.dnz + 1 + CVR_BRICK - 1) / CVR_BRICK;
    size_t n_bricks = (size_t)h->maj_dim[0] * h->maj_dim[1] * h->maj_dim[2];
    CVR_CUDA(h, vol_alloc(h, (void**)&h->d_majorant, n_bricks * sizeof(float)));
    k_build_majorant<<<(unsigned)((n_bricks * 32 + bt - 1) / bt), bt, 0, h->stream>>>(
        (const float4*)h->d_dcells, m.dnx, m.dny, m.dnz, h->maj_dim[0], h->maj_dim[1], h->maj_dim[2], h->d_majorant);
    CVR_CUDA(h, cudaGetLastError());
    if (build_majorant2(h)) return 1;
  }
  pt.mark("density alloc + copy + build");
  m.albedo_const = s->albedo ? 0 : 1;
  m.albedo_r = s->albedo_const[0], m.albedo_g = s->albedo_const[1], m.albedo_b = s->albedo_const[2];
  m.anx = m.any = m.anz = 2;
  if (s->albedo) {
    m.anx = s->albedo_dim[0], m.any = s->albedo_dim[1], m.anz = s->albedo_dim[2];
    size_t na = (size_t)m.anx * m.any * m.anz;
    CVR_CUDA(h, vol_alloc(h, (void**)&h->d_albedo, na * sizeof(float4)));
    CVR_CUDA(h, cudaMemcpyAsync(h->d_albedo, s->albedo, na * sizeof(float4), kind, h->stream));
    if (h->layout == LAYOUT_CELL8) {
      size_t ncell = (size_t)(m.anx + 1) * (m.any + 1) * (m.anz + 1);
      CVR_CUDA(h, vol_alloc(h, (void**)&h->d_acells, ncell * CVR_ACELL_FLOATS * sizeof(float)));
      int g = (int)std::min<size_t>((ncell * 8 + bt - 1) / bt, (size_t)h->sm_count * 32);
      k_build_albedo_cells<<<g, bt, 0, h->stream>>>(h->d_albedo, m.anx, m.any, m.anz, h->d_acells);
      CVR_CUDA(h, cudaGetLastError());
    }
  }
  pt.mark("albedo alloc + copy + build");
  if (h->layout == LAYOUT_CELL8) {
    // the dense copies are only the source of the cell layouts
    CVR_CUDA(h, cudaStreamSynchronize(h->stream));
    pt.mark("sync 2");
    vol_free(h, h->d_density);
    vol_free(h, h->d_albedo);
    h->d_density = nullptr;
    h->d_albedo = nullptr;
  }
  else {
    // layout=linear: the H2D copies above are asynchronous for pinned host buffers, and the host
    // pointers are borrowed for the duration of this call only
    CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  pt.mark("free staging");
  m.density = h->d_density;
  m.dcells = h->d_dcells;
  m.albedo = h->d_albedo;
  m.acells = h->d_acells;
  {  // device footprint of the density lookups (selects the warp scheduler's slot count)
    const int before = effective_wslots(h);
    h->volume_bytes = h->layout == LAYOUT_CELL8 ? (size_t)(m.dnx + 1) * (m.dny + 1) * (m.dnz + 1) * 32 : nd * sizeof(float);
    if (effective_wslots(h) != before) h->inited = false;  // another kernel instantiation
  }
  h->scene_set = true;
  h->skip_dirty = true;  // new majorants: the fetch-skip table is planned and built again at the next launch
  h->inited = false;
  return 0;
}

extern "C++" {
namespace {

// medium scalars shared by the sparse / procedural scene setters
void set_medium_scalars(cvr_handle h, const float box_min[3], const float box_max[3], float scale, float hg_g,
                        const float ggx_alpha[2], float ggx_eta, const float albedo_const[3]) {
  MediumParams& m = h->P.med;
  m.box_min = V3{box_min[0], box_min[1], box_min[2]};
  m.box_max = V3{box_max[0], box_max[1], box_max[2]};
  m.scale = scale;
  m.hg_g = hg_g;
  m.alpha_x = ggx_alpha && ggx_alpha[0] > 0.f ? ggx_alpha[0] : 0.1f;
  m.alpha_y = ggx_alpha && ggx_alpha[1] > 0.f ? ggx_alpha[1] : 0.1f;
  m.eta = ggx_eta > 0.f ? ggx_eta : 1.05f / 1.01f;
  m.albedo_const = 1;
  m.albedo_r = albedo_const[0], m.albedo_g = albedo_const[1], m.albedo_b = albedo_const[2];
  m.anx = m.any = m.anz = 2;
  m.density = nullptr, m.albedo = nullptr, m.acells = nullptr;
}

// bricks + slot table + majorant grid from a voxel accessor; returns the maximum voxel value
template <class Acc>
int build_bricks(cvr_handle h, const Acc& acc, int nx, int ny, int nz, float* max_value) {
  MediumParams& m = h->P.med;
  m.dnx = nx, m.dny = ny, m.dnz = nz;
  const uint32_t bmx = (uint32_t)(nx + 1 + 7) / 8, bmy = (uint32_t)(ny + 1 + 7) / 8, bmz = (uint32_t)(nz + 1 + 7) / 8;
  const size_t nb = (size_t)bmx * bmy * bmz;
  if (nb >= (1ull << 31)) return fail(h, "sparse scene: brick grid too large");
  uint32_t* d_counter = nullptr;
  CVR_CUDA(h, vol_alloc(h, (void**)&h->d_btable, nb * sizeof(uint32_t)));
  CVR_CUDA(h, cudaMalloc(&d_counter, 2 * sizeof(uint32_t)));
  CVR_CUDA(h, cudaMemsetAsync(d_counter, 0, 2 * sizeof(uint32_t), h->stream));
  const int bt = 256;
  const int g = (int)std::min<size_t>((nb + bt - 1) / bt, (size_t)h->sm_count * 32);
  k_brick_slots_fn<<<g, bt, 0, h->stream>>>(acc, nx, ny, nz, bmx, bmy, bmz, h->d_btable, d_counter);
  CVR_CUDA(h, cudaGetLastError());
  uint32_t n_slots = 0;
  CVR_CUDA(h, cudaMemcpyAsync(&n_slots, d_counter, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_bricks = n_slots;
  const size_t brick_floats = 512 * 8;
  cudaError_t e = vol_alloc(h, (void**)&h->d_dcells, ((size_t)n_slots + 1) * brick_floats * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(d_counter);
    return fail(h, "sparse scene: %u bricks (%.1f GB) do not fit: %s", n_slots,
                ((double)n_slots + 1) * brick_floats * 4 / 1e9, cudaGetErrorString(e));
  }
  CVR_CUDA(h, cudaMemsetAsync(h->d_dcells, 0, brick_floats * sizeof(float), h->stream));  // slot 0 = the zero brick
  // one CTA per brick-grid entry: grid.x = one z-slice of the brick grid, grid.y = slices
  k_build_bricks_fn<<<dim3(bmx * bmy, bmz), 512, 0, h->stream>>>(acc, nx, ny, nz, bmx, bmy, h->d_btable,
                                                                (float4*)h->d_dcells, n_slots);
  CVR_CUDA(h, cudaGetLastError());
  h->maj_dim[0] = bmx, h->maj_dim[1] = bmy, h->maj_dim[2] = bmz;
  CVR_CUDA(h, vol_alloc(h, (void**)&h->d_majorant, nb * sizeof(float)));
  k_brick_majorant<<<(unsigned)((nb * 32 + bt - 1) / bt), bt, 0, h->stream>>>(h->d_btable, (const float4*)h->d_dcells, nb,
                                                                                h->d_majorant, d_counter + 1);
  CVR_CUDA(h, cudaGetLastError());
  if (build_majorant2(h)) return 1;
  uint32_t max_bits = 0;
  CVR_CUDA(h, cudaMemcpyAsync(&max_bits, d_counter + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_counter);
  std::memcpy(max_value, &max_bits, 4);
  m.dcells = h->d_dcells;
  m.btable = h->d_btable;
  m.bmx = bmx, m.bmy = bmy, m.bmz = bmz;
  h->scene_layout = LAYOUT_BRICK;
  h->volume_bytes = ((size_t)n_slots + 1) * brick_floats * sizeof(float) + nb * sizeof(uint32_t);
  h->inited = false;
  return 0;
}

}  // namespace
}  // extern "C++"

int cvr_set_scene_sparse(cvr_handle h, const cvr_sparse_desc* s) {
  CVR_CHECK_HANDLE(h);
  if (!s || !s->leaf_origins || !s->leaf_values || s->n_leaves == 0) return fail(h, "cvr_set_scene_sparse: no leaves");
  if (s->dim[0] < 2 || s->dim[1] < 2 || s->dim[2] < 2) return fail(h, "cvr_set_scene_sparse: grid dims must be >= 2");
  if (!(s->scale > 0.f)) return fail(h, "cvr_set_scene_sparse: scale must be positive");
  if (set_device(h)) return 1;
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  free_volume(h);
  set_medium_scalars(h, s->box_min, s->box_max, s->scale, s->hg_g, s->ggx_alpha, s->ggx_eta, s->albedo_const);
  // leaf grid: absolute index space from the multiple of 8 at or below bbox_min
  int base[3], off[3], lb[3];
  for (int a = 0; a < 3; ++a) {
    base[a] = (int)std::floor(s->bbox_min[a] / 8.0) * 8;
    off[a] = s->bbox_min[a] - base[a];
    lb[a] = (off[a] + s->dim[a] + 7) / 8;
  }
  const size_t n_table = (size_t)lb[0] * lb[1] * lb[2];
  int32_t* d_org = nullptr;
  float* d_val = nullptr;
  uint32_t* d_tab = nullptr;
  CVR_CUDA(h, vol_alloc(h, (void**)&d_org, s->n_leaves * 3 * sizeof(int32_t)));
  CVR_CUDA(h, vol_alloc(h, (void**)&d_val, s->n_leaves * 512 * sizeof(float)));
  CVR_CUDA(h, vol_alloc(h, (void**)&d_tab, n_table * sizeof(uint32_t)));
  CVR_CUDA(h, cudaMemcpyAsync(d_org, s->leaf_origins, s->n_leaves * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  CVR_CUDA(h, cudaMemcpyAsync(d_val, s->leaf_values, s->n_leaves * 512 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  CVR_CUDA(h, cudaMemsetAsync(d_tab, 0, n_table * sizeof(uint32_t), h->stream));
  k_fill_leaf_table<<<(unsigned)((s->n_leaves + 255) / 256), 256, 0, h->stream>>>(d_org, s->n_leaves, base[0], base[1], base[2],
                                                                                   lb[0], lb[1], lb[2], d_tab);
  CVR_CUDA(h, cudaGetLastError());
  LeafVolumeAccessor acc{LeafAccessor{d_tab, d_val, lb[0], lb[1], lb[2], off[0], off[1], off[2]}};
  float mx = 0.f;
  int rc = build_bricks(h, acc, s->dim[0], s->dim[1], s->dim[2], &mx);
  vol_free(h, d_org), vol_free(h, d_val), vol_free(h, d_tab);
  if (rc) return rc;
  h->P.med.max_density = s->max_density > 0.f ? s->max_density : mx;  // VDBSceneBuilder.h:54-55: max voxel
  if (!(h->P.med.max_density > 0.f)) return fail(h, "cvr_set_scene_sparse: the volume is empty");
  h->scene_set = true;
  h->skip_dirty = true;  // new majorants: the fetch-skip table is planned and built again at the next launch
  h->inited = false;
  return 0;
}

int cvr_set_scene_procedural(cvr_handle h, const char* kind_c, int32_t n, uint32_t seed, const cvr_scene_desc* s,
                             float* max_density_out) {
  CVR_CHECK_HANDLE(h);
  if (!kind_c || !s) return fail(h, "cvr_set_scene_procedural: null argument");
  const std::string kind(kind_c);
  if (kind != "fbm" && kind != "sparsefbm") return fail(h, "cvr_set_scene_procedural: unknown kind '%s' (fbm | sparsefbm)", kind_c);
  if (n < 8 || n > 4096) return fail(h, "cvr_set_scene_procedural: n must be in [8, 4096]");
  if (!(s->scale > 0.f)) return fail(h, "cvr_set_scene_procedural: scale must be positive");
  if (set_device(h)) return 1;
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  free_volume(h);
  set_medium_scalars(h, s->box_min, s->box_max, s->scale, s->hg_g, s->ggx_alpha, s->ggx_eta, s->albedo_const);
  const FbmAccessor acc{seed ? seed : 0x5eedu, kind == "sparsefbm" ? 1 : 0};
  float mx = 0.f;
  if (kind == "sparsefbm") {
    if (int rc = build_bricks(h, acc, n, n, n, &mx)) return rc;
  } else {  // dense cell8 generated in place
    MediumParams& m = h->P.med;
    m.dnx = m.dny = m.dnz = n;
    m.btable = nullptr, m.bmx = m.bmy = m.bmz = 0;
    const size_t ncell = (size_t)(n + 1) * (n + 1) * (n + 1);
    if (ncell >= (1ull << 32)) return fail(h, "cvr_set_scene_procedural: grid too large for 32-bit cell indices");
    CVR_CUDA(h, vol_alloc(h, (void**)&h->d_dcells, ncell * 8 * sizeof(float)));
    const int bt = 256;
    const int g = (int)std::min<size_t>((ncell + bt - 1) / bt, (size_t)h->sm_count * 32);
    k_build_cells_fn<<<g, bt, 0, h->stream>>>(acc, n, n, n, (float4*)h->d_dcells);
    CVR_CUDA(h, cudaGetLastError());
    h->maj_dim[0] = h->maj_dim[1] = h->maj_dim[2] = (uint32_t)(n + 1 + CVR_BRICK - 1) / CVR_BRICK;
    const size_t nbk = (size_t)h->maj_dim[0] * h->maj_dim[1] * h->maj_dim[2];
    CVR_CUDA(h, vol_alloc(h, (void**)&h->d_majorant, nbk * sizeof(float)));
    k_build_majorant<<<(unsigned)((nbk * 32 + bt - 1) / bt), bt, 0, h->stream>>>(
        (const float4*)h->d_dcells, n, n, n, h->maj_dim[0], h->maj_dim[1], h->maj_dim[2], h->d_majorant);
    CVR_CUDA(h, cudaGetLastError());
    if (build_majorant2(h)) return 1;
    // max voxel = max over the majorant grid
    std::vector<float> maj(nbk);
    CVR_CUDA(h, cudaMemcpyAsync(maj.data(), h->d_majorant, nbk * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CVR_CUDA(h, cudaStreamSynchronize(h->stream));
    for (float v : maj) mx = std::max(mx, v);
    m.dcells = h->d_dcells;
    h->scene_layout = LAYOUT_CELL8;
    h->volume_bytes = ncell * 32;
    h->inited = false;
  }
  h->P.med.max_density = s->max_density > 0.f ? s->max_density : (mx > 0.f ? mx : 1.f);
  if (max_density_out) *max_density_out = h->P.med.max_density;
  h->scene_set = true;
  h->skip_dirty = true;  // new majorants: the fetch-skip table is planned and built again at the next launch
  h->inited = false;
  return 0;
}

int cvr_get_volume_info(cvr_handle h, uint64_t* layout_bytes, uint64_t* n_bricks, int32_t* layout) {
  CVR_CHECK_HANDLE(h);
  if (layout_bytes) *layout_bytes = h->volume_bytes;
  if (n_bricks) *n_bricks = h->n_bricks;
  if (layout) *layout = h->scene_layout;
  return 0;
}

int cvr_set_resolution(cvr_handle h, uint32_t tile_w, uint32_t tile_h) {
  CVR_CHECK_HANDLE(h);
  if (tile_w == 0 || tile_h == 0) return fail(h, "cvr_set_resolution: zero tile");
  h->tile_w = tile_w, h->tile_h = tile_h;
  h->P.cam.res_x = (float)tile_w;  // make_float2(resolution.x, resolution.y), RenderKernelLauncher.h:35
  h->P.cam.res_y = (float)tile_h;
  return 0;
}

int cvr_set_pixel_index_range(cvr_handle h, float full_w, float full_h) {
  CVR_CHECK_HANDLE(h);
  h->P.cam.range_x = full_w, h->P.cam.range_y = full_h;
  return 0;
}

int cvr_set_raster_to_view(cvr_handle h, float x, float y) {
  CVR_CHECK_HANDLE(h);
  h->P.cam.rtv_x = x, h->P.cam.rtv_y = y;
  return 0;
}

int cvr_set_inv_view_matrix(cvr_handle h, const float m[12]) {
  CVR_CHECK_HANDLE(h);
  if (!m) return fail(h, "cvr_set_inv_view_matrix: null");
  memcpy(h->P.cam.m, m, 12 * sizeof(float));
  return 0;
}

int cvr_set_offset(cvr_handle h, uint32_t x, uint32_t y) {
  CVR_CHECK_HANDLE(h);
  h->P.off_x = x, h->P.off_y = y;
  return 0;
}

int cvr_set_output(cvr_handle h, void* d_output_float4) {
  CVR_CHECK_HANDLE(h);
  h->d_out = (float4*)d_output_float4;
  return 0;
}

int cvr_set_iterations(cvr_handle h, uint32_t n) {
  CVR_CHECK_HANDLE(h);
  if (n == 0) return fail(h, "cvr_set_iterations: zero iterations");
  h->iterations = n;
  return 0;
}

int cvr_get_iterations(cvr_handle h, uint32_t* n) {
  CVR_CHECK_HANDLE(h);
  if (n) *n = h->iterations;
  return 0;
}

int cvr_set_seed(cvr_handle h, uint32_t seed) {
  CVR_CHECK_HANDLE(h);
  h->seed = seed;
  return 0;
}

int cvr_get_seed(cvr_handle h, uint32_t* seed) {
  CVR_CHECK_HANDLE(h);
  if (seed) *seed = h->seed;
  return 0;
}

int cvr_set_sample_range(cvr_handle h, uint32_t first, uint32_t count) {
  CVR_CHECK_HANDLE(h);
  // (0, 0) = every sample.  The range is checked against the iteration count when a render is
  // launched (the setters may come in any order): check_sample_range
  h->sample_first = first, h->sample_count = count;
  return 0;
}

int cvr_init(cvr_handle h) {
  CVR_CHECK_HANDLE(h);
  if (set_device(h)) return 1;
  h->inited = false;
  return ensure_init(h);
}

int cvr_allocate(cvr_handle h) {
  CVR_CHECK_HANDLE(h);
  if (set_device(h)) return 1;
  return ensure_allocated(h);
}

int cvr_launch_render(cvr_handle h) {
  CVR_CHECK_HANDLE(h);
  if (!h->d_out) return fail(h, "cvr_launch_render before cvr_set_output");
  if (check_sample_range(h)) return 1;
  unsigned long long b, e;
  path_range(h, b, e);
  // naiveSK seeds with the bare path id (NaiveVolPTsk_kernel.cuh:22, Q7)
  uint32_t seed = h->variant == VAR_NAIVE ? 0u : h->seed;
  return launch(h, h->d_out, h->tile_w, 0, nullptr, 1, 0, 0, seed, 0, nullptr, b, e);
}

int cvr_reset(cvr_handle h) {
  CVR_CHECK_HANDLE(h);
  if (set_device(h)) return 1;
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  // RenderKernelLauncher.cu:353-361 (regeneration: seed_ += n_paths_), :567-575 (streaming: seed_++)
  uint32_t n_paths = (uint32_t)(h->P.cam.res_x * h->P.cam.res_y) * h->iterations;
  if (h->variant == VAR_REGEN || h->variant == VAR_STREAM_MK) h->seed += n_paths;
  if (h->variant == VAR_STREAM) h->seed += 1;
  return 0;
}

int cvr_sync(cvr_handle h) {
  CVR_CHECK_HANDLE(h);
  if (set_device(h)) return 1;
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  return 0;
}

int cvr_get_counters(cvr_handle h, cvr_counters* out) {
  CVR_CHECK_HANDLE(h);
  if (!out) return fail(h, "cvr_get_counters: null");
  if (set_device(h)) return 1;
  memset(out, 0, sizeof *out);
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  collect_timing(h);
  if (h->allocated) {
    DeviceCounters c;
    CVR_CUDA(h, cudaMemcpy(&c, h->d_ctr, sizeof c, cudaMemcpyDeviceToHost));
    out->paths = c.paths, out->bounces = c.bounces, out->density_lookups = c.density_lookups;
    out->albedo_lookups = c.albedo_lookups, out->escaped = c.escaped;
    out->speculative_lookups = c.speculative;
    out->skipped_fetches = c.skipped;
  }
  out->launches = h->launches;
  out->kernel_ms = h->kernel_ms;
  return 0;
}

int cvr_reset_counters(cvr_handle h) {
  CVR_CHECK_HANDLE(h);
  if (set_device(h)) return 1;
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  collect_timing(h);
  if (h->allocated) CVR_CUDA(h, cudaMemset(h->d_ctr, 0, sizeof(DeviceCounters)));
  h->launches = 0;
  h->kernel_ms = 0.0;
  return 0;
}

int cvr_get_launch_shape(cvr_handle h, int* grid, int* block, int* regs) {
  CVR_CHECK_HANDLE(h);
  if (set_device(h) || ensure_init(h)) return 1;
  if (grid) *grid = h->grid;
  if (block) *block = effective_block(h);
  if (regs) *regs = h->regs;
  return 0;
}

int cvr_resolve_tile(cvr_handle h, const void* d_tile, uint32_t tile_w, uint32_t tile_h, void* d_image,
                     uint32_t full_w, uint32_t full_h, uint32_t off_x, uint32_t off_y, float scale) {
  CVR_CHECK_HANDLE(h);
  if (!d_tile || !d_image) return fail(h, "cvr_resolve_tile: null buffer");
  if (off_x + tile_w > full_w || off_y + tile_h > full_h) return fail(h, "cvr_resolve_tile: tile outside image");
  if (set_device(h)) return 1;
  uint32_t n = tile_w * tile_h;
  int g = (int)std::min<uint32_t>((n + 255) / 256, (uint32_t)h->sm_count * 8);
  k_resolve_tile<<<g, 256, 0, h->stream>>>((const float4*)d_tile, tile_w, tile_h, tile_w, 0, 0,
                                           (float4*)d_image, full_w, off_x, off_y, scale);
  CVR_CUDA(h, cudaGetLastError());
  return 0;
}

int cvr_resolve_tile_display(cvr_handle h, const void* d_tile, uint32_t tile_w, uint32_t tile_h, void* d_transfer,
                             void* d_display, uint32_t full_w, uint32_t full_h, uint32_t off_x, uint32_t off_y, float scale,
                             int reset_transfer) {
  CVR_CHECK_HANDLE(h);
  if (!d_tile || !d_transfer || !d_display) return fail(h, "cvr_resolve_tile_display: null buffer");
  if (!tile_w || !tile_h || !full_w || !full_h) return fail(h, "cvr_resolve_tile_display: zero size");
  if (!(scale > 0.f)) return fail(h, "cvr_resolve_tile_display: scale must be positive");
  if (set_device(h)) return 1;
  // ImageBufferTransfer.cu:143-147: the transfer buffer starts from zero with the first iteration
  if (reset_transfer) CVR_CUDA(h, cudaMemsetAsync(d_transfer, 0, (size_t)full_w * full_h * sizeof(float4), h->stream));
  const uint32_t n = tile_w * tile_h;
  const int g = (int)std::min<uint32_t>((n + 255) / 256, (uint32_t)h->sm_count * 8);
  k_accumulate_display<<<g, 256, 0, h->stream>>>((const float4*)d_tile, tile_w, tile_h, (float4*)d_transfer, (uchar4*)d_display,
                                                 full_w, full_h, off_x, off_y, scale);
  CVR_CUDA(h, cudaGetLastError());
  return 0;
}

int cvr_tile_table(uint32_t res_x, uint32_t res_y, uint32_t ntx, uint32_t nty, uint32_t tile_dim[2],
                   uint32_t* origins) {
  if (!ntx || !nty || !tile_dim) return 1;
  // Config.h:70-71: ceil() of an INTEGER division, i.e. floor (Q6)
  tile_dim[0] = (uint32_t)(int)std::ceil((double)(res_x / ntx));
  tile_dim[1] = (uint32_t)(int)std::ceil((double)(res_y / nty));
  if (origins) {
    for (uint32_t id = 0; id < ntx * nty; ++id) {
      // CudaVolPath.cpp:15-19
      origins[2 * id + 0] = tile_dim[0] * (id % ntx);
      origins[2 * id + 1] = tile_dim[1] * (uint32_t)(int)((float)id / (float)ntx);
    }
  }
  return 0;
}

int cvr_default_camera(uint32_t res_x, uint32_t res_y, float fov_x, float inv_view[12],
                       float raster_to_view[2]) {
  // Camera.h:25-37 (right (1,0,0), up (0,-1,0), view (0,0,-1), position (0,0,100)) laid out
  // as CudaVolPath.cpp:71-84 does: three rows, translation in the fourth column
  static const float m[12] = {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 100.0f};
  if (inv_view) memcpy(inv_view, m, sizeof m);
  if (raster_to_view) {
    float fov_y = ((float)res_y / (float)res_x) * fov_x;  // Camera.h:63-67
    raster_to_view[0] = tanf(fov_x * CVR_PI / 360.f);     // Camera.h:69-71
    raster_to_view[1] = tanf(fov_y * CVR_PI / 360.f);
  }
  return 0;
}

extern "C++" {
namespace {

// One phase of a render: tiles k = tile_first, tile_first + tile_stride, ... < tile_limit of the
// reference's tile list, each with sample indices [sample_first, sample_first + sample_count)
// (count kAllSamples = up to the iteration count; 0 = nothing).  Phases of one call own disjoint tiles.
constexpr uint32_t kAllSamples = 0xffffffffu;
struct RenderPhase {
  uint32_t tile_first, tile_stride, tile_limit, sample_first, sample_count;
};
// the phase's sample range against the iteration count: false = nothing to render
bool phase_samples(cvr_handle h, const RenderPhase& ph, uint32_t iterations) {
  if (ph.sample_first >= iterations) return false;
  const uint32_t cnt = ph.sample_count == kAllSamples ? iterations - ph.sample_first : ph.sample_count;
  if (!cnt) return false;
  h->sample_first = ph.sample_first, h->sample_count = cnt;
  return true;
}

bool phase_owns(const RenderPhase& ph, uint32_t k) {
  return k >= ph.tile_first && k < ph.tile_limit && (k - ph.tile_first) % ph.tile_stride == 0;
}

// CudaVolPath::render (CudaVolPath.cpp:338-347) over the given phases.  `zero_image`: clear the
// resolved image first (multi-GPU: every rank contributes a full-resolution image to a sum).
// `add_into_image`: the resolve ADDS into d_image_out (which may be another device's memory) instead of storing (device groups).
int render_phases(cvr_handle h, const cvr_render_desc* r, const RenderPhase* phases, int n_phases, float* host_image,
                  void* d_image_out, bool zero_image, bool sync_at_end, bool add_into_image = false) {
  if (!r) return fail(h, "cvr_render_image: null desc");
  if (!r->res_x || !r->res_y || !r->n_tiles_x || !r->n_tiles_y || !r->iterations)
    return fail(h, "cvr_render_image: zero resolution / tiles / iterations");
  if (r->n_tiles_x > r->res_x || r->n_tiles_y > r->res_y) return fail(h, "cvr_render_image: more tiles than pixels");
  if (!h->scene_set) return fail(h, "cvr_render_image before cvr_set_scene");
  if (set_device(h)) return 1;
  PhaseTimer pt("cvr_render_image");
  const uint32_t n_tiles = r->n_tiles_x * r->n_tiles_y;
  std::vector<uint32_t> origins(2 * (size_t)n_tiles);
  uint32_t tile_dim[2];
  cvr_tile_table(r->res_x, r->res_y, r->n_tiles_x, r->n_tiles_y, tile_dim, origins.data());
  float inv_view[12], rtv[2];
  cvr_default_camera(r->res_x, r->res_y, r->fov_x > 0.f ? r->fov_x : 0.7f, inv_view, rtv);
  if (r->inv_view) memcpy(inv_view, r->inv_view, sizeof inv_view);
  if (r->raster_to_view) memcpy(rtv, r->raster_to_view, sizeof rtv);
  // constructor + render() prologue order of CudaVolPath (CudaVolPath.cpp:39-58, 338-341)
  cvr_set_raster_to_view(h, rtv[0], rtv[1]);
  cvr_set_resolution(h, tile_dim[0], tile_dim[1]);
  cvr_set_pixel_index_range(h, (float)r->res_x, (float)r->res_y);
  if (ensure_init(h) || ensure_allocated(h)) return 1;
  cvr_set_iterations(h, r->iterations);
  cvr_set_inv_view_matrix(h, inv_view);

  const size_t image_px = (size_t)r->res_x * r->res_y;
  const size_t tile_px = (size_t)tile_dim[0] * tile_dim[1];
  if (h->d_image_px < image_px) {
    if (h->d_image) cudaFreeAsync(h->d_image, h->stream);
    h->d_image = nullptr;
    h->d_image_px = 0;
    CVR_CUDA(h, cudaMallocAsync((void**)&h->d_image, image_px * sizeof(float4), h->stream));
    h->d_image_px = image_px;
  }
  float4* d_image = d_image_out ? (float4*)d_image_out : h->d_image;
  if (zero_image) CVR_CUDA(h, cudaMemsetAsync(d_image, 0, image_px * sizeof(float4), h->stream));
  const uint32_t n_paths_tile = (uint32_t)tile_px * r->iterations;  // uint, RenderKernelLauncher.cu:125
  // stream base of global tile k = seed0 + k * step, i.e. what reset() accumulates on one GPU
  const uint32_t seed0 = h->variant == VAR_NAIVE ? 0u : h->seed;
  const uint32_t seed_step = (h->variant == VAR_REGEN || h->variant == VAR_STREAM_MK) ? n_paths_tile : h->variant == VAR_STREAM ? 1u : 0u;
  const float scale = (float)r->iterations;  // UtilityFunctors::Scale(current_iteration_)
  const int rg = (int)std::min<size_t>((tile_px + 255) / 256, (size_t)h->sm_count * 8);
  auto owned = [&](uint32_t k) {
    for (int i = 0; i < n_phases; ++i)
      if (phase_owns(phases[i], k)) return true;
    return false;
  };
  for (int i = 0; i < n_phases; ++i) {
    if (!phases[i].tile_stride) return fail(h, "cvr_render_image: zero tile stride");
    for (int j = 0; j < i; ++j)
      for (uint32_t k = 0; k < n_tiles; ++k)
        if (phase_owns(phases[i], k) && phase_owns(phases[j], k)) return fail(h, "cvr_render_image: phases share tile %u", k);
  }

  if (add_into_image && !(r->fuse_tiles && n_phases <= 2 && n_tiles <= 65535u))
    return fail(h, "cvr_render_image: the adding resolve needs fuse_tiles, at most two phases and at most 65535 tiles");
  if (r->fuse_tiles) {
    // one launch per phase over every tile it owns, accumulating straight into a full-resolution
    // buffer; same pixels, same streams as the loop below
    if (h->d_tile_px < image_px) {
      if (h->d_tile) cudaFreeAsync(h->d_tile, h->stream);
      h->d_tile = nullptr;
      h->d_tile_px = 0;
      CVR_CUDA(h, cudaMallocAsync((void**)&h->d_tile, image_px * sizeof(float4), h->stream));
      h->d_tile_px = image_px;
    }
    if (h->d_origins_n < n_tiles) {
      if (h->d_origins) cudaFreeAsync(h->d_origins, h->stream);
      h->d_origins = nullptr;
      h->d_origins_n = 0;
      CVR_CUDA(h, cudaMallocAsync((void**)&h->d_origins, n_tiles * sizeof(uint2), h->stream));
      h->d_origins_n = n_tiles;
    }
    CVR_CUDA(h, cudaMemcpyAsync(h->d_origins, origins.data(), n_tiles * sizeof(uint2),
                                cudaMemcpyHostToDevice, h->stream));
    CVR_CUDA(h, cudaMemsetAsync(h->d_tile, 0, image_px * sizeof(float4), h->stream));
    for (int i = 0; i < n_phases; ++i) {
      const RenderPhase& ph = phases[i];
      const uint32_t limit = std::min(ph.tile_limit, n_tiles);
      const uint32_t n_mine = ph.tile_first < limit ? (limit - ph.tile_first + ph.tile_stride - 1) / ph.tile_stride : 0;
      if (!n_mine || !phase_samples(h, ph, r->iterations)) continue;
      if (check_sample_range(h)) return 1;
      unsigned long long pb, pe;
      path_range(h, pb, pe);
      if (pe == pb) continue;
      if (launch(h, h->d_tile, r->res_x, 1, h->d_origins, n_mine, ph.tile_first, ph.tile_stride, seed0, seed_step, nullptr, pb, pe))
        return 1;
    }
    if (n_phases <= 2 && n_tiles <= 65535u) {
      ResolveOwner own{};
      own.n = (uint32_t)n_phases;
      for (int i = 0; i < n_phases; ++i)
        own.first[i] = phases[i].tile_first, own.stride[i] = phases[i].tile_stride, own.limit[i] = std::min(phases[i].tile_limit, n_tiles);
      const unsigned bx = (unsigned)std::max<size_t>(1, std::min<size_t>((tile_px + 255) / 256, ((size_t)h->sm_count * 8 + n_tiles - 1) / n_tiles));
      if (add_into_image)
        k_resolve_tiles<true><<<dim3(bx, n_tiles), 256, 0, h->stream>>>(h->d_tile, tile_dim[0], tile_dim[1], r->res_x,
                                                                         (const uint2*)h->d_origins, d_image, scale, own);
      else
        k_resolve_tiles<false><<<dim3(bx, n_tiles), 256, 0, h->stream>>>(h->d_tile, tile_dim[0], tile_dim[1], r->res_x,
                                                                          (const uint2*)h->d_origins, d_image, scale, own);
    } else {
      for (uint32_t k = 0; k < n_tiles; ++k) {
        if (!owned(k)) continue;
        uint32_t ox = origins[2 * k], oy = origins[2 * k + 1];
        k_resolve_tile<<<rg, 256, 0, h->stream>>>(h->d_tile, tile_dim[0], tile_dim[1], r->res_x, ox, oy,
                                                  d_image, r->res_x, ox, oy, scale);
      }
    }
    CVR_CUDA(h, cudaGetLastError());
  } else {
    if (h->d_tile_px < tile_px) {
      if (h->d_tile) cudaFreeAsync(h->d_tile, h->stream);
      h->d_tile = nullptr;
      h->d_tile_px = 0;
      CVR_CUDA(h, cudaMallocAsync((void**)&h->d_tile, tile_px * sizeof(float4), h->stream));
      h->d_tile_px = tile_px;
    }
    for (int i = 0; i < n_phases; ++i) {
      const RenderPhase& ph = phases[i];
      unsigned long long pb = 0, pe = 0;
      if (phase_samples(h, ph, r->iterations)) {
        if (check_sample_range(h)) return 1;
        path_range(h, pb, pe);
      }
      for (uint32_t k = ph.tile_first; k < std::min(ph.tile_limit, n_tiles); k += ph.tile_stride) {
        uint32_t ox = origins[2 * k], oy = origins[2 * k + 1];
        // initRenderState / prepareForNextIterations: zeroed tile buffer (CudaVolPath.cpp:188-208)
        CVR_CUDA(h, cudaMemsetAsync(h->d_tile, 0, tile_px * sizeof(float4), h->stream));
        cvr_set_offset(h, ox, oy);  // copyOffset (CudaVolPath.cpp:260)
        if (pe > pb && launch(h, h->d_tile, tile_dim[0], 0, nullptr, 1, 0, 0, seed0 + k * seed_step, 0, nullptr, pb, pe)) return 1;
        k_resolve_tile<<<rg, 256, 0, h->stream>>>(h->d_tile, tile_dim[0], tile_dim[1], tile_dim[0], 0, 0,
                                                  d_image, r->res_x, ox, oy, scale);
        CVR_CUDA(h, cudaGetLastError());
      }
    }
  }
  cvr_set_sample_range(h, 0, 0);
  pt.mark("setup + launches (host side)");
  if (host_image) {
    for (uint32_t k = 0; k < n_tiles; ++k) {
      if (!owned(k)) continue;
      uint32_t ox = origins[2 * k], oy = origins[2 * k + 1];
      size_t off = (size_t)oy * r->res_x + ox;
      CVR_CUDA(h, cudaMemcpy2DAsync(host_image + 4 * off, (size_t)r->res_x * sizeof(float4), d_image + off,
                                    (size_t)r->res_x * sizeof(float4), tile_dim[0] * sizeof(float4),
                                    tile_dim[1], cudaMemcpyDeviceToHost, h->stream));
    }
  }
  pt.mark("D2H enqueue");
  if (sync_at_end) CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  pt.mark("final sync (kernel + copies)");
  // what the sequence of reset() calls leaves behind after all tiles
  h->seed += n_tiles * seed_step;
  return 0;
}

}  // namespace
}  // extern "C++"

int cvr_render_image(cvr_handle h, const cvr_render_desc* r, float* host_image, void* d_image_out) {
  CVR_CHECK_HANDLE(h);
  if (!r) return fail(h, "cvr_render_image: null desc");
  if (r->sample_first && r->sample_first >= r->iterations)
    return fail(h, "cvr_render_image: first sample %u is not below the iteration count %u", r->sample_first, r->iterations);
  if (r->sample_count && (unsigned long long)r->sample_first + r->sample_count > r->iterations)
    return fail(h, "cvr_render_image: samples [%u, %llu) exceed the iteration count %u", r->sample_first,
                (unsigned long long)r->sample_first + r->sample_count, r->iterations);
  const RenderPhase ph{r->tile_first, r->tile_stride ? r->tile_stride : 1u, 0xffffffffu, r->sample_first,
                       r->sample_count ? r->sample_count : kAllSamples};
  return render_phases(h, r, &ph, 1, host_image, d_image_out, false, true);
}

/* Balanced static sharding (100 tiles over 8 ranks = 13 vs 12):
 *   mode 0 "tiles"     rank r renders tiles k = r (mod world), all samples
 *   mode 1 "spp"       every rank renders every tile, samples [first, first + count)
 *   mode 2 "balanced"  the complete rounds of the interleave (tiles < world * floor(n_tiles / world))
 *                      go by tile, the left-over tiles are split by SAMPLE INDEX over all ranks:
 *                      every rank gets n_tiles / world tile-equivalents of work, every (tile,
 *                      sample) pair is rendered exactly once, with the stream it has on one GPU. */
int cvr_shard_plan(uint32_t n_tiles, uint32_t iterations, int rank, int world, int mode, cvr_shard* out) {
  if (!out || world < 1 || rank < 0 || rank >= world || !n_tiles || !iterations || mode < 0 || mode > 2) return 1;
  memset(out, 0, sizeof *out);
  const uint32_t w = (uint32_t)world, r = (uint32_t)rank;
  // sample split of `iterations` over the ranks: remainders go to the low ranks
  const uint32_t base = iterations / w, rem = iterations % w;
  const uint32_t s_first = r * base + std::min(r, rem), s_count = base + (r < rem ? 1u : 0u);
  if (mode == 0) {
    out->tile_first = r, out->tile_stride = w, out->tile_limit = n_tiles;
    out->tail_first = out->tail_limit = n_tiles;
  } else if (mode == 1) {
    out->tile_first = out->tile_limit = 0, out->tile_stride = 1;  // no whole tiles
    out->tail_first = 0, out->tail_limit = n_tiles;
    out->sample_first = s_first, out->sample_count = s_count;
  } else if (rem == 0) {
    // the samples divide evenly: the sample split IS the balanced plan -- every rank traces exactly the same number of
    // paths through the same pixels in ONE launch.  Measured, C3 on 8 ranks, per-rank kernel ms (tools/shard_balance.py,
    // ideal 10.46): sample split 10.86-10.88, tile interleave 10.52-10.95 (tiles are unequal work), rounds of tiles +
    // left-over tiles by sample 10.73-11.15 (two launches = two drain tails of ~0.4 ms).
    out->tile_first = out->tile_limit = 0, out->tile_stride = 1;
    out->tail_first = 0, out->tail_limit = n_tiles;
    out->sample_first = s_first, out->sample_count = s_count;
  } else {
    const uint32_t whole = (n_tiles / w) * w;
    out->tile_first = r, out->tile_stride = w, out->tile_limit = whole;
    out->tail_first = whole, out->tail_limit = n_tiles;
    out->sample_first = s_first, out->sample_count = s_count;
  }
  return 0;
}

int cvr_render_image_sharded(cvr_handle h, const cvr_render_desc* r, const cvr_shard* sh, float* host_image, void* d_image_out) {
  CVR_CHECK_HANDLE(h);
  if (!r || !sh) return fail(h, "cvr_render_image_sharded: null argument");
  RenderPhase ph[2];
  int n = 0;
  if (sh->tile_first < sh->tile_limit)
    ph[n++] = RenderPhase{sh->tile_first, sh->tile_stride ? sh->tile_stride : 1u, sh->tile_limit, 0u, kAllSamples};
  // a rank whose sample share of the tail is empty (more ranks than samples) renders nothing of it
  if (sh->tail_first < sh->tail_limit && sh->sample_count)
    ph[n++] = RenderPhase{sh->tail_first, 1u, sh->tail_limit, sh->sample_first, sh->sample_count};
  // every rank's image is a term of a sum over ranks: pixels it does not own must be zero
  return render_phases(h, r, ph, n, host_image, d_image_out, true, true);
}

// =========================================================================== device groups
}  // extern "C" (the group implementation needs C++ linkage for its helpers)

#include <dlfcn.h>
#include <nccl.h>  // types and enums only: the library is loaded with dlopen on first use

#include <map>
#include <mutex>
#include <thread>

struct cvr_group {
  std::vector<cvr_handle> members;
  std::vector<int> devices;
  std::vector<ncclComm_t> comms;     // empty until the first reduce of a group with > 1 device
  std::vector<float4*> d_images;     // per-member full-resolution resolved image (group-owned)
  size_t image_px = 0;
  // "peer" reduce: every device adds its resolved share straight into d_peer_image on the first device through NVLink
  // peer memory (k_resolve_tiles<true>); ordered by events instead of a collective.  Needs peer access from every
  // member to the first; otherwise (or with the option group_reduce=nccl) the shares are summed by ONE ncclReduce.
  bool peer_ok = false;
  int reduce_mode = -1;              // -1 auto (= nccl: measured, below), 0 nccl, 1 peer
  float4* d_peer_image = nullptr;    // plain cudaMalloc on the first device (pool memory is not peer-mapped by default)
  size_t peer_image_px = 0;
  cudaEvent_t ev_zero = nullptr;
  std::vector<cudaEvent_t> ev_done;
  std::string err;
};

namespace {

thread_local std::string g_group_create_error;

int gfail(cvr_group_handle g, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (g)
    g->err = buf;
  else
    g_group_create_error = buf;
  return 1;
}

// the few NCCL entry points a framebuffer reduce needs, resolved from libnccl.so.2 at run time
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) {
    api.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
    return api;
  }
#define CVR_NCCL_SYM(field, sym)                                                  \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym));         \
  if (!api.field) {                                                               \
    api.why = std::string("libnccl has no symbol ") + sym;                        \
    api.lib = nullptr;                                                            \
    return api;                                                                   \
  }
  CVR_NCCL_SYM(CommInitAll, "ncclCommInitAll")
  CVR_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  CVR_NCCL_SYM(GroupStart, "ncclGroupStart")
  CVR_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  CVR_NCCL_SYM(Reduce, "ncclReduce")
  CVR_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CVR_NCCL_SYM
  return api;
}

// ncclCommInitAll over 8 devices takes ~1.1 s: communicators outlive their group and are handed to the next group over
// the same device list (cvr_render builds a renderer per trial, Main.cpp runTest; measured before this cache: 1.20 s
// "rendering time" per trial on 8 GPUs against 0.108 s on one -- all of it communicator set-up inside the first reduce)
std::mutex g_comm_cache_mutex;
std::map<std::vector<int>, std::vector<std::vector<ncclComm_t>>> g_comm_cache;

int group_comms(cvr_group_handle g) {
  if (!g->comms.empty()) return 0;
  NcclApi& N = nccl_api();
  if (!N.lib) return gfail(g, "cvr_group: %s (groups of more than one device need NCCL)", N.why.c_str());
  {
    std::lock_guard<std::mutex> lock(g_comm_cache_mutex);
    auto it = g_comm_cache.find(g->devices);
    if (it != g_comm_cache.end() && !it->second.empty()) {
      g->comms = std::move(it->second.back());
      it->second.pop_back();
      return 0;
    }
  }
  g->comms.assign(g->members.size(), nullptr);
  ncclResult_t rc = N.CommInitAll(g->comms.data(), (int)g->members.size(), g->devices.data());
  if (rc != ncclSuccess) {
    g->comms.clear();
    return gfail(g, "ncclCommInitAll failed: %s", N.GetErrorString(rc));
  }
  return 0;
}

// run f(rank) on one host thread per member; returns the first failing rank + 1 (0 = all fine)
template <class F>
int for_each_member(cvr_group_handle g, F f) {
  const int n = (int)g->members.size();
  std::vector<int> rc((size_t)n, 0);
  if (n == 1) {
    rc[0] = f(0);
  } else {
    std::vector<std::thread> th;
    for (int r = 0; r < n; ++r) th.emplace_back([&, r]() { rc[(size_t)r] = f(r); });
    for (auto& t : th) t.join();
  }
  for (int r = 0; r < n; ++r)
    if (rc[(size_t)r]) return r + 1;
  return 0;
}

int member_failed(cvr_group_handle g, int bad, const char* what) {
  return gfail(g, "%s: device %d: %s", what, g->devices[(size_t)bad - 1], cvr_last_error(g->members[(size_t)bad - 1]));
}

}  // namespace

extern "C" {

const char* cvr_group_last_error(cvr_group_handle g) { return g ? g->err.c_str() : g_group_create_error.c_str(); }

int cvr_group_create(const char* kernel_name, const int* devices, int n_devices, cvr_group_handle* out) {
  if (!out) return gfail(nullptr, "cvr_group_create: out is null");
  *out = nullptr;
  if (n_devices < 1) return gfail(nullptr, "cvr_group_create: need at least one device");
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return gfail(nullptr, "cvr_group_create: no CUDA device (%s); there is no CPU fallback", cudaGetErrorString(e));
  cvr_group_handle g = new cvr_group();
  for (int r = 0; r < n_devices; ++r) {
    const int dev = devices ? devices[r] : r;
    for (int q : g->devices)
      if (q == dev) {
        cvr_group_destroy(g);
        return gfail(nullptr, "cvr_group_create: device %d listed twice", dev);
      }
    cvr_handle h = nullptr;
    if (cvr_create(kernel_name, dev, &h)) {
      std::string msg = cvr_last_error(nullptr);
      cvr_group_destroy(g);
      return gfail(nullptr, "cvr_group_create: %s", msg.c_str());
    }
    g->members.push_back(h);
    g->devices.push_back(dev);
    g->d_images.push_back(nullptr);
  }
  if (n_devices > 1) {
    // peer access member -> first device (NVLink / NVSwitch): lets a member's resolve add into the first device's image
    bool ok = true;
    for (int r = 1; r < n_devices && ok; ++r) {
      int can = 0;
      ok = cudaDeviceCanAccessPeer(&can, g->devices[(size_t)r], g->devices[0]) == cudaSuccess && can;
      if (ok) {
        cudaSetDevice(g->devices[(size_t)r]);
        cudaError_t pe = cudaDeviceEnablePeerAccess(g->devices[0], 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError(), pe = cudaSuccess;
        ok = pe == cudaSuccess;
      }
    }
    (void)cudaGetLastError();
    g->peer_ok = ok;
    g->ev_done.assign((size_t)n_devices, nullptr);
    for (int r = 0; r < n_devices; ++r) {
      cudaSetDevice(g->devices[(size_t)r]);
      cudaEventCreateWithFlags(&g->ev_done[(size_t)r], cudaEventDisableTiming);
    }
    cudaSetDevice(g->devices[0]);
    cudaEventCreateWithFlags(&g->ev_zero, cudaEventDisableTiming);
  }
  if (n_devices > 1 && group_comms(g)) {  // set-up cost and a missing NCCL belong to create, not to the first render
    std::string msg = g->err;
    cvr_group_destroy(g);
    return gfail(nullptr, "cvr_group_create: %s", msg.c_str());
  }
  *out = g;
  return 0;
}

int cvr_group_destroy(cvr_group_handle g) {
  if (!g) return 0;
  if (!g->comms.empty()) {
    // quiesce, then park the communicators for the next group over the same devices (group_comms)
    for (size_t r = 0; r < g->members.size(); ++r) {
      cudaSetDevice(g->devices[r]);
      cudaStreamSynchronize(g->members[r]->stream);
    }
    std::lock_guard<std::mutex> lock(g_comm_cache_mutex);
    g_comm_cache[g->devices].push_back(std::move(g->comms));
  }
  if (!g->devices.empty()) {
    cudaSetDevice(g->devices[0]);
    if (g->d_peer_image) cudaFree(g->d_peer_image);
    if (g->ev_zero) cudaEventDestroy(g->ev_zero);
  }
  for (size_t r = 0; r < g->ev_done.size(); ++r)
    if (g->ev_done[r]) {
      cudaSetDevice(g->devices[r]);
      cudaEventDestroy(g->ev_done[r]);
    }
  for (size_t r = 0; r < g->members.size(); ++r) {
    if (g->d_images[r]) {
      cudaSetDevice(g->devices[r]);
      cudaFreeAsync(g->d_images[r], g->members[r]->stream);
    }
    cvr_destroy(g->members[r]);
  }
  delete g;
  return 0;
}

int cvr_group_size(cvr_group_handle g, int* n) {
  if (!g) return gfail(nullptr, "null group");
  if (n) *n = (int)g->members.size();
  return 0;
}

int cvr_group_member(cvr_group_handle g, int rank, cvr_handle* member) {
  if (!g) return gfail(nullptr, "null group");
  if (rank < 0 || rank >= (int)g->members.size() || !member) return gfail(g, "cvr_group_member: rank %d of %zu", rank, g->members.size());
  *member = g->members[(size_t)rank];
  return 0;
}

int cvr_group_set_option(cvr_group_handle g, const char* key, const char* value) {
  if (!g) return gfail(nullptr, "null group");
  if (key && value && std::string(key) == "group_reduce") {  // how the members' shares are summed: auto | peer | nccl
    const std::string v(value);
    if (v != "auto" && v != "peer" && v != "nccl") return gfail(g, "group_reduce: unknown value '%s' (auto | peer | nccl)", value);
    if (v == "peer" && g->members.size() > 1 && !g->peer_ok)
      return gfail(g, "group_reduce=peer: device %d has no peer access to device %d", g->devices.back(), g->devices[0]);
    g->reduce_mode = v == "auto" ? -1 : v == "peer" ? 1 : 0;
    return 0;
  }
  for (size_t r = 0; r < g->members.size(); ++r)
    if (cvr_set_option(g->members[r], key, value)) return member_failed(g, (int)r + 1, "cvr_group_set_option");
  return 0;
}

int cvr_group_set_seed(cvr_group_handle g, uint32_t seed) {
  if (!g) return gfail(nullptr, "null group");
  for (cvr_handle h : g->members) cvr_set_seed(h, seed);
  return 0;
}

int cvr_group_set_scene(cvr_group_handle g, const cvr_scene_desc* scene) {
  if (!g) return gfail(nullptr, "null group");
  if (int bad = for_each_member(g, [&](int r) { return cvr_set_scene(g->members[(size_t)r], scene); }))
    return member_failed(g, bad, "cvr_group_set_scene");
  return 0;
}

int cvr_group_set_scene_sparse(cvr_group_handle g, const cvr_sparse_desc* scene) {
  if (!g) return gfail(nullptr, "null group");
  if (int bad = for_each_member(g, [&](int r) { return cvr_set_scene_sparse(g->members[(size_t)r], scene); }))
    return member_failed(g, bad, "cvr_group_set_scene_sparse");
  return 0;
}

int cvr_group_set_scene_procedural(cvr_group_handle g, const char* kind, int32_t n, uint32_t seed, const cvr_scene_desc* medium,
                                   float* max_density_out) {
  if (!g) return gfail(nullptr, "null group");
  std::vector<float> mx(g->members.size(), 0.f);
  if (int bad = for_each_member(g, [&](int r) {
        return cvr_set_scene_procedural(g->members[(size_t)r], kind, n, seed, medium, &mx[(size_t)r]);
      }))
    return member_failed(g, bad, "cvr_group_set_scene_procedural");
  if (max_density_out) *max_density_out = mx[0];
  return 0;
}

int cvr_group_reduce(cvr_group_handle g, void* const* d_buffers, uint64_t n_floats) {
  if (!g) return gfail(nullptr, "null group");
  if (!d_buffers) return gfail(g, "cvr_group_reduce: null buffer list");
  const int n = (int)g->members.size();
  if (n > 1) {
    if (group_comms(g)) return 1;
    NcclApi& N = nccl_api();
    ncclResult_t rc = N.GroupStart();
    for (int r = 0; r < n && rc == ncclSuccess; ++r) {
      if (!d_buffers[r]) {
        N.GroupEnd();
        return gfail(g, "cvr_group_reduce: buffer %d is null", r);
      }
      cudaSetDevice(g->devices[(size_t)r]);
      // in place on the root; recvbuff is only significant there
      rc = N.Reduce(d_buffers[r], d_buffers[r], (size_t)n_floats, ncclFloat32, ncclSum, 0, g->comms[(size_t)r],
                    g->members[(size_t)r]->stream);
    }
    ncclResult_t rc2 = N.GroupEnd();
    if (rc != ncclSuccess || rc2 != ncclSuccess)
      return gfail(g, "ncclReduce failed: %s", N.GetErrorString(rc != ncclSuccess ? rc : rc2));
  }
  for (int r = 0; r < n; ++r) {
    cudaSetDevice(g->devices[(size_t)r]);
    cudaError_t e = cudaStreamSynchronize(g->members[(size_t)r]->stream);
    if (e != cudaSuccess) return gfail(g, "cvr_group_reduce: device %d: %s", g->devices[(size_t)r], cudaGetErrorString(e));
  }
  return 0;
}

int cvr_group_render_image(cvr_group_handle g, const cvr_render_desc* desc, int shard_mode, float* host_image, void* d_image_rank0_out) {
  if (!g) return gfail(nullptr, "null group");
  if (!desc || !desc->res_x || !desc->res_y || !desc->n_tiles_x || !desc->n_tiles_y || !desc->iterations)
    return gfail(g, "cvr_group_render_image: zero resolution / tiles / iterations");
  const int n = (int)g->members.size();
  const uint32_t n_tiles = desc->n_tiles_x * desc->n_tiles_y;
  std::vector<cvr_shard> plan((size_t)n);
  for (int r = 0; r < n; ++r)
    if (cvr_shard_plan(n_tiles, desc->iterations, r, n, shard_mode, &plan[(size_t)r]))
      return gfail(g, "cvr_group_render_image: bad shard mode %d", shard_mode);
  const size_t image_px = (size_t)desc->res_x * desc->res_y;
  auto phases_of = [&](int r, RenderPhase* ph) {
    int np = 0;
    const cvr_shard& sh = plan[(size_t)r];
    if (sh.tile_first < sh.tile_limit)
      ph[np++] = RenderPhase{sh.tile_first, sh.tile_stride ? sh.tile_stride : 1u, sh.tile_limit, 0u, kAllSamples};
    if (sh.tail_first < sh.tail_limit && sh.sample_count)
      ph[np++] = RenderPhase{sh.tail_first, 1u, sh.tail_limit, sh.sample_first, sh.sample_count};
    return np;
  };
  PhaseTimer pt("cvr_group_render_image");
  std::vector<void*> bufs((size_t)n);
  // Measured on 2 x B200 (NV18), C3 1024^2 x 256 spp (gpurun call AA): peer 42.5 ms + 1.4 ms of enqueue, nccl 42.2 + 0.6 + 0.6 --
  // a 16 MiB ncclReduce costs < 0.1 ms next to a 42 ms (84 / N) render, and cudaMalloc of a peer-mapped image is slow
  // (outliers of 50 ms when a renderer is built per trial).  The fused form is therefore opt-in, not the default.
  const bool peer = n > 1 && g->peer_ok && g->reduce_mode == 1 && desc->fuse_tiles && n_tiles <= 65535u;
  if (n > 1 && g->reduce_mode == 1 && !peer)
    return gfail(g, "cvr_group_render_image: group_reduce=peer needs peer access and fuse_tiles");
  if (peer) {
    // ---- resolve fused with the sum: every member adds its share into ONE image on the first device (peer memory)
    cudaSetDevice(g->devices[0]);
    if (g->peer_image_px < image_px) {
      cudaStreamSynchronize(g->members[0]->stream);
      cudaFree(g->d_peer_image);
      g->d_peer_image = nullptr, g->peer_image_px = 0;
      cudaError_t e = cudaMalloc((void**)&g->d_peer_image, image_px * sizeof(float4));
      if (e != cudaSuccess) return gfail(g, "cvr_group_render_image: device %d: %s", g->devices[0], cudaGetErrorString(e));
      g->peer_image_px = image_px;
    }
    for (int r = 0; r < n; ++r) bufs[(size_t)r] = g->d_peer_image;
    cudaError_t e = cudaMemsetAsync(g->d_peer_image, 0, image_px * sizeof(float4), g->members[0]->stream);
    if (e == cudaSuccess) e = cudaEventRecord(g->ev_zero, g->members[0]->stream);
    if (e != cudaSuccess) return gfail(g, "cvr_group_render_image: device %d: %s", g->devices[0], cudaGetErrorString(e));
    if (int bad = for_each_member(g, [&](int r) {
          cvr_handle m = g->members[(size_t)r];
          if (set_device(m)) return 1;
          // nobody adds before the image is zero; the adds are ordered behind this member's own render by its stream
          if (r > 0 && cudaStreamWaitEvent(m->stream, g->ev_zero, 0) != cudaSuccess) return fail(m, "cudaStreamWaitEvent failed");
          RenderPhase ph[2];
          const int np = phases_of(r, ph);
          if (render_phases(m, desc, ph, np, nullptr, g->d_peer_image, false, false, true)) return 1;
          if (cudaEventRecord(g->ev_done[(size_t)r], m->stream) != cudaSuccess) return fail(m, "cudaEventRecord failed");
          return 0;
        }))
      return member_failed(g, bad, "cvr_group_render_image");
    pt.mark("renders enqueued");
    cudaSetDevice(g->devices[0]);
    for (int r = 1; r < n; ++r)
      if (cudaStreamWaitEvent(g->members[0]->stream, g->ev_done[(size_t)r], 0) != cudaSuccess)
        return gfail(g, "cvr_group_render_image: cudaStreamWaitEvent failed");
  } else {
    if (g->image_px < image_px) {
      for (int r = 0; r < n; ++r) {
        cudaSetDevice(g->devices[(size_t)r]);
        if (g->d_images[(size_t)r]) cudaFreeAsync(g->d_images[(size_t)r], g->members[(size_t)r]->stream);
        g->d_images[(size_t)r] = nullptr;
        cudaError_t e = cudaMallocAsync((void**)&g->d_images[(size_t)r], image_px * sizeof(float4), g->members[(size_t)r]->stream);
        if (e != cudaSuccess) {
          g->image_px = 0;
          return gfail(g, "cvr_group_render_image: device %d: %s", g->devices[(size_t)r], cudaGetErrorString(e));
        }
      }
      g->image_px = image_px;
    }
    for (int r = 0; r < n; ++r) bufs[(size_t)r] = g->d_images[(size_t)r];
    if (d_image_rank0_out) bufs[0] = d_image_rank0_out;
    // every rank renders its share into its own zeroed full-resolution image (one host thread each)
    if (int bad = for_each_member(g, [&](int r) {
          RenderPhase ph[2];
          const int np = phases_of(r, ph);
          // no sync here: the reduce is queued behind the render on the member's stream
          return render_phases(g->members[(size_t)r], desc, ph, np, nullptr, bufs[(size_t)r], true, false);
        }))
      return member_failed(g, bad, "cvr_group_render_image");
    pt.mark("renders enqueued");
    if (n > 1 && cvr_group_reduce(g, bufs.data(), (uint64_t)image_px * 4)) return 1;
  }
  if (n > 1) {
    if (!peer) pt.mark("renders + reduce done");
    // alpha is "some path escaped" / iterations, not a sum over the ranks that saw one (k_clamp_alpha)
    cudaSetDevice(g->devices[0]);
    k_clamp_alpha<<<g->members[0]->sm_count * 4, 256, 0, g->members[0]->stream>>>((float4*)bufs[0], image_px,
                                                                                   1.0f / (float)desc->iterations);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && peer && d_image_rank0_out)  // the caller's buffer need not be peer-mapped: hand the sum over on device 0
      e = cudaMemcpyAsync(d_image_rank0_out, g->d_peer_image, image_px * sizeof(float4), cudaMemcpyDeviceToDevice, g->members[0]->stream);
    if (e != cudaSuccess) return gfail(g, "cvr_group_render_image: alpha: %s", cudaGetErrorString(e));
  }
  if (host_image) {
    // the covered region (tile_dim * n_tiles; Q6: remainder pixels are never rendered and left untouched)
    uint32_t tile_dim[2];
    cvr_tile_table(desc->res_x, desc->res_y, desc->n_tiles_x, desc->n_tiles_y, tile_dim, nullptr);
    cudaSetDevice(g->devices[0]);
    cudaError_t e = cudaMemcpy2DAsync(host_image, (size_t)desc->res_x * sizeof(float4), bufs[0], (size_t)desc->res_x * sizeof(float4),
                                      (size_t)tile_dim[0] * desc->n_tiles_x * sizeof(float4), (size_t)tile_dim[1] * desc->n_tiles_y,
                                      cudaMemcpyDeviceToHost, g->members[0]->stream);
    if (e != cudaSuccess) return gfail(g, "cvr_group_render_image: copy to host: %s", cudaGetErrorString(e));
  }
  for (int r = 0; r < n; ++r) {
    cudaSetDevice(g->devices[(size_t)r]);
    cudaError_t e = cudaStreamSynchronize(g->members[(size_t)r]->stream);
    if (e != cudaSuccess) return gfail(g, "cvr_group_render_image: device %d: %s", g->devices[(size_t)r], cudaGetErrorString(e));
  }
  pt.mark("image on the host, streams idle");
  return 0;
}

int cvr_group_get_counters(cvr_group_handle g, cvr_counters* out) {
  if (!g) return gfail(nullptr, "null group");
  if (!out) return gfail(g, "cvr_group_get_counters: null");
  memset(out, 0, sizeof *out);
  for (size_t r = 0; r < g->members.size(); ++r) {
    cvr_counters c;
    if (cvr_get_counters(g->members[r], &c)) return member_failed(g, (int)r + 1, "cvr_group_get_counters");
    out->paths += c.paths, out->bounces += c.bounces, out->density_lookups += c.density_lookups;
    out->albedo_lookups += c.albedo_lookups, out->escaped += c.escaped, out->speculative_lookups += c.speculative_lookups;
    out->launches += c.launches, out->skipped_fetches += c.skipped_fetches;
    out->kernel_ms = std::max(out->kernel_ms, c.kernel_ms);
  }
  return 0;
}

int cvr_group_reset_counters(cvr_group_handle g) {
  if (!g) return gfail(nullptr, "null group");
  for (size_t r = 0; r < g->members.size(); ++r)
    if (cvr_reset_counters(g->members[r])) return member_failed(g, (int)r + 1, "cvr_group_reset_counters");
  return 0;
}

int cvr_trace_paths(cvr_handle h, uint64_t first, uint64_t count, void* d_per_path) {
  CVR_CHECK_HANDLE(h);
  if (!d_per_path || !count) return fail(h, "cvr_trace_paths: null buffer / zero count");
  if (set_device(h)) return 1;
  CVR_CUDA(h, cudaMemsetAsync(d_per_path, 0, count * sizeof(float4), h->stream));
  uint32_t seed = h->variant == VAR_NAIVE ? 0u : h->seed;
  return launch(h, nullptr, h->tile_w, 0, nullptr, 1, 0, 0, seed, 0, (float4*)d_per_path, first,
                first + count);
}

int cvr_trace_paths_logged(cvr_handle h, uint64_t first, uint64_t count, void* d_per_path, void* d_log, uint32_t log_cap) {
  CVR_CHECK_HANDLE(h);
  if (!d_per_path || !count) return fail(h, "cvr_trace_paths_logged: null buffer / zero count");
  if (!d_log || !log_cap) return fail(h, "cvr_trace_paths_logged: null log / zero capacity");
  if (h->rng_mode != RNG_XORWOW_PATH || h->sched != 3 || h->tracking != 0)
    return fail(h, "cvr_trace_paths_logged: needs rng=xorwow-path, sched=warp, tracking=global");
  if (set_device(h)) return 1;
  CVR_CUDA(h, cudaMemsetAsync(d_per_path, 0, count * sizeof(float4), h->stream));
  CVR_CUDA(h, cudaMemsetAsync(d_log, 0, count * (size_t)log_cap * sizeof(uint2), h->stream));
  h->trace_log = (uint2*)d_log, h->trace_log_cap = log_cap;
  uint32_t seed = h->variant == VAR_NAIVE ? 0u : h->seed;
  int rc = launch(h, nullptr, h->tile_w, 0, nullptr, 1, 0, 0, seed, 0, (float4*)d_per_path, first, first + count);
  h->trace_log = nullptr, h->trace_log_cap = 0;
  return rc;
}

int cvr_rng_kat(cvr_handle h, const int32_t* seeds, int n_seeds, int n, uint32_t* words, float* uniforms) {
  CVR_CHECK_HANDLE(h);
  if (!seeds || n_seeds <= 0 || n <= 0 || !words || !uniforms) return fail(h, "cvr_rng_kat: bad arguments");
  if (set_device(h)) return 1;
  int32_t* d_s;
  uint32_t* d_w;
  float* d_u;
  size_t tot = (size_t)n_seeds * n;
  CVR_CUDA(h, cudaMalloc(&d_s, n_seeds * sizeof(int32_t)));
  CVR_CUDA(h, cudaMalloc(&d_w, tot * 4));
  CVR_CUDA(h, cudaMalloc(&d_u, tot * 4));
  CVR_CUDA(h, cudaMemcpyAsync(d_s, seeds, n_seeds * sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
  k_rng_kat<<<(n_seeds + 63) / 64, 64, 0, h->stream>>>(d_s, n_seeds, n, d_w, d_u);
  CVR_CUDA(h, cudaGetLastError());
  CVR_CUDA(h, cudaMemcpyAsync(words, d_w, tot * 4, cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaMemcpyAsync(uniforms, d_u, tot * 4, cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_s), cudaFree(d_w), cudaFree(d_u);
  return 0;
}

int cvr_debug_trig_check(cvr_handle h, float limit, uint64_t mismatches[3], uint32_t first_bad_bits[3]) {
  CVR_CHECK_HANDLE(h);
  if (!mismatches || !(limit > 0.f) || !(limit < 105615.0f)) return fail(h, "cvr_debug_trig_check: bad arguments (0 < limit < 105615)");
  if (set_device(h)) return 1;
  unsigned long long* d_m;
  uint32_t* d_f;
  CVR_CUDA(h, cudaMalloc(&d_m, 3 * sizeof(unsigned long long)));
  CVR_CUDA(h, cudaMalloc(&d_f, 3 * sizeof(uint32_t)));
  CVR_CUDA(h, cudaMemsetAsync(d_m, 0, 3 * sizeof(unsigned long long), h->stream));
  CVR_CUDA(h, cudaMemsetAsync(d_f, 0xff, 3 * sizeof(uint32_t), h->stream));
  uint32_t limit_bits;
  std::memcpy(&limit_bits, &limit, 4);
  k_trig_check<<<h->sm_count * 16, 256, 0, h->stream>>>(limit_bits, d_m, d_f);
  CVR_CUDA(h, cudaGetLastError());
  unsigned long long m[3];
  uint32_t f[3];
  CVR_CUDA(h, cudaMemcpyAsync(m, d_m, sizeof m, cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaMemcpyAsync(f, d_f, sizeof f, cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_m), cudaFree(d_f);
  for (int i = 0; i < 3; ++i) {
    mismatches[i] = m[i];
    if (first_bad_bits) first_bad_bits[i] = f[i];
  }
  return 0;
}

int cvr_gather_roofline(cvr_handle h, uint64_t footprint_bytes, int loads_per_thread, int unroll, double* gbs) {
  CVR_CHECK_HANDLE(h);
  if (!gbs || footprint_bytes < 4096 || loads_per_thread < 1) return fail(h, "cvr_gather_roofline: bad arguments");
  if (set_device(h)) return 1;
  uint64_t n_cells = footprint_bytes / 32;
  if (n_cells >= (1ull << 32)) n_cells = (1ull << 32) - 1;
  float *d_cells, *d_sink;
  CVR_CUDA(h, cudaMalloc(&d_cells, n_cells * 32));
  CVR_CUDA(h, cudaMalloc(&d_sink, 4));
  CVR_CUDA(h, cudaMemsetAsync(d_cells, 0, n_cells * 32, h->stream));
  const int grid = h->sm_count * 8, block = 256;
  loads_per_thread = (loads_per_thread + 7) / 8 * 8;
  cudaEvent_t e0, e1;
  CVR_CUDA(h, cudaEventCreate(&e0));
  CVR_CUDA(h, cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition warms the caches / TLB
    CVR_CUDA(h, cudaEventRecord(e0, h->stream));
    if (unroll < 0)  // negative unroll: 8 loads in flight, L1 bypassed (ld.global.nc.L1::no_allocate)
      k_gather_bench<8, true><<<grid, block, 0, h->stream>>>(d_cells, (uint32_t)n_cells, loads_per_thread, d_sink);
    else if (unroll >= 8)
      k_gather_bench<8><<<grid, block, 0, h->stream>>>(d_cells, (uint32_t)n_cells, loads_per_thread, d_sink);
    else if (unroll >= 4)
      k_gather_bench<4><<<grid, block, 0, h->stream>>>(d_cells, (uint32_t)n_cells, loads_per_thread, d_sink);
    else
      k_gather_bench<1><<<grid, block, 0, h->stream>>>(d_cells, (uint32_t)n_cells, loads_per_thread, d_sink);
    CVR_CUDA(h, cudaGetLastError());
    CVR_CUDA(h, cudaEventRecord(e1, h->stream));
    CVR_CUDA(h, cudaEventSynchronize(e1));
    float ms = 0.f;
    CVR_CUDA(h, cudaEventElapsedTime(&ms, e0, e1));
    double v = (double)grid * block * loads_per_thread * 32.0 / (ms * 1e-3) / 1e9;
    if (rep > 0 && v > best) best = v;
  }
  cudaEventDestroy(e0), cudaEventDestroy(e1);
  cudaFree(d_cells), cudaFree(d_sink);
  *gbs = best;
  return 0;
}

int cvr_debug_lookup(cvr_handle h, const float* p, int n, float* dens, float* alb) {
  CVR_CHECK_HANDLE(h);
  if (!h->scene_set) return fail(h, "cvr_debug_lookup before cvr_set_scene");
  if (!p || n <= 0 || !dens || !alb) return fail(h, "cvr_debug_lookup: bad arguments");
  if (set_device(h)) return 1;
  fill_track_inv(h->P);
  float *d_p, *d_d, *d_a;
  CVR_CUDA(h, cudaMalloc(&d_p, (size_t)n * 12));
  CVR_CUDA(h, cudaMalloc(&d_d, (size_t)n * 4));
  CVR_CUDA(h, cudaMalloc(&d_a, (size_t)n * 12));
  CVR_CUDA(h, cudaMemcpyAsync(d_p, p, (size_t)n * 12, cudaMemcpyHostToDevice, h->stream));
  if (h->scene_layout == LAYOUT_BRICK)
    k_debug_lookup<LAYOUT_BRICK><<<(n + 127) / 128, 128, 0, h->stream>>>(h->P.med, h->P.inv, d_p, n, d_d, d_a);
  else if (h->scene_layout == LAYOUT_CELL8)
    k_debug_lookup<LAYOUT_CELL8><<<(n + 127) / 128, 128, 0, h->stream>>>(h->P.med, h->P.inv, d_p, n, d_d, d_a);
  else
    k_debug_lookup<LAYOUT_LINEAR><<<(n + 127) / 128, 128, 0, h->stream>>>(h->P.med, h->P.inv, d_p, n, d_d, d_a);
  CVR_CUDA(h, cudaGetLastError());
  CVR_CUDA(h, cudaMemcpyAsync(dens, d_d, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaMemcpyAsync(alb, d_a, (size_t)n * 12, cudaMemcpyDeviceToHost, h->stream));
  CVR_CUDA(h, cudaStreamSynchronize(h->stream));
  cudaFree(d_p), cudaFree(d_d), cudaFree(d_a);
  return 0;
}

}  // extern "C"
