// cvr_kernels.cuh -- the sm_100a path kernels.
//
// One persistent-thread kernel template replaces the reference's naiveSK /
// regenerationSK(thread) / streamingSK kernels (NaiveVolPTsk_kernel.cuh:17-87,
// RegenerationVolPTsk_kernel.cuh:146-232, StreamingVolPTsk_kernel.cuh:219-290):
// the kernel NAME selects the reference SEMANTICS (scatter pull-back, seed
// advance, roulette-after-escape), the scheduling is always the B200 one:
//
//  * grid = SMs x resident CTAs, every warp loops until the path queue is empty;
//  * idle lanes claim path ids with ONE warp-aggregated 64-bit atomic
//    (ballot + popc + shuffle), consecutive ids -> consecutive pixels;
//  * a per-lane state machine splits each bounce into {intersect, Woodcock steps,
//    event}; the Woodcock loop runs warp-wide ("while-while") and yields to the
//    event phase once `loop_threshold` lanes are waiting, so a lane's RNG draw
//    order is exactly the reference's while divergence is bounded;
//  * density lookups read one 32-byte cell (8 trilinear corners) with a single
//    256-bit load; albedo cells are one 128-byte line.
#pragma once
#include "cvr_device.cuh"

namespace cvr {

enum : int { RNG_XORWOW_PATH = 0, RNG_XORWOW_THREAD = 1, RNG_PHILOX = 2 };
enum : int { LAYOUT_LINEAR = 0, LAYOUT_CELL8 = 1 };

struct DeviceCounters {
  unsigned long long paths, bounces, density_lookups, albedo_lookups, escaped, speculative;
};

struct KernelParams {
  CameraParams cam;
  MediumParams med;
  // work: this launch runs, for each of n_launch_tiles tiles, the path ids
  // [path_begin, path_end) of that tile (path id = sample * npix + pixel).
  unsigned long long path_begin, path_end;
  uint32_t npix;    // (uint)(c_resolution.x * c_resolution.y)
  uint32_t tile_w;  // (uint)c_resolution.x
  uint32_t off_x, off_y;          // c_offset when tile_origins == nullptr
  const uint2* tile_origins;      // fused-tile mode: origin of global tile k
  uint32_t n_launch_tiles;        // tiles covered by this launch (1 in single-tile mode)
  uint32_t tile_first, tile_stride;  // global tile index of launch tile j = first + j*stride
  uint32_t seed;                  // stream base of global tile 0
  uint32_t seed_step;             // added per global tile index (fused mode)
  // output
  float4* out;          // accumulation buffer
  uint32_t out_stride;  // row stride in pixels
  int out_full;         // 0: index by tile-local pixel, 1: by full-image pixel (fused)
  float4* per_path;     // debug: per-path radiance (or nullptr)
  unsigned long long* head;  // path queue head
  DeviceCounters* ctr;
  uint32_t max_bounces;
  int loop_threshold;
  int pullback;         // naive/streaming: o -= d*EPSILON at scatter (Q8)
  int rr;               // Russian roulette enabled (Defines.h:44)
  int rr_after_escape;  // thread-rng regeneration/streaming: the roulette draw is consumed after an escape (Q8)
};

enum : int { S_IDLE = 0, S_ISECT = 1, S_TRACK = 2, S_SCATTER = 3, S_BOUNDARY = 4, S_DONE = 5 };

CVR_DEV V3 normal_from_code(int c) {
  // 0:+x 1:+y 2:+z 3:-x 4:-y 5:-z
  float s = c < 3 ? 1.0f : -1.0f;
  int a = c < 3 ? c : c - 3;
  return v3(a == 0 ? s : 0.0f, a == 1 ? s : 0.0f, a == 2 ? s : 0.0f);
}
CVR_DEV int code_from_normal(V3 n) {
  if (n.x > 0.5f) return 0;
  if (n.y > 0.5f) return 1;
  if (n.z > 0.5f) return 2;
  if (n.x < -0.5f) return 3;
  if (n.y < -0.5f) return 4;
  return 5;
}

template <int LAYOUT>
CVR_DEV float density_lookup(const MediumParams& m, V3 p) {
  if (LAYOUT == LAYOUT_CELL8) return density_cell8(m, p);
  return density_linear(m, p);
}
template <int LAYOUT>
CVR_DEV V3 albedo_lookup(const MediumParams& m, V3 p) {
  if (m.albedo_const) return v3(m.albedo_r, m.albedo_g, m.albedo_b);
  if (LAYOUT == LAYOUT_CELL8) return albedo_cell8(m, p);
  return albedo_linear(m, p);
}

template <int RNGM>
struct RngSel {
  typedef Xorwow type;
};
template <>
struct RngSel<RNG_PHILOX> {
  typedef Philox type;
};

#ifndef CVR_BLOCK
#define CVR_BLOCK 256
#endif
#ifndef CVR_MIN_BLOCKS
#define CVR_MIN_BLOCKS 4
#endif

template <int RNGM, int LAYOUT, bool COUNT>
__global__ void __launch_bounds__(CVR_BLOCK, CVR_MIN_BLOCKS)
    k_volpt(const __grid_constant__ KernelParams P) {
  typedef typename RngSel<RNGM>::type Rng;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned lane_lt = (1u << lane) - 1u;

  Rng rng;
  if (RNGM == RNG_XORWOW_THREAD) {
    // RegenerationVolPTsk_kernel.cuh:151,156: Rng rng(seed + tid), once per thread
    uint32_t tid = threadIdx.x + blockDim.x * blockIdx.x;
    rng.init((int32_t)(P.seed + tid));
  }

  const unsigned long long per_tile = P.path_end - P.path_begin;
  const unsigned long long total = per_tile * P.n_launch_tiles;

  // loop invariants of Woodcock tracking (Utilities.cuh:143, 129-132)
  const float inv_max_sigmat = 1.0f / (P.med.scale * P.med.max_density);
  const V3 q = P.med.box_min / (P.med.box_max - P.med.box_min);

  int state = S_IDLE;
  bool exhausted = false;
  V3 o = v3(0, 0, 0), d = v3(0, 0, 1);
  float thr_x = 1.f, thr_y = 1.f, thr_z = 1.f;
  float t = 0.f, dist = 0.f;
  int ncode = 0;
  uint32_t out_idx = 0, bounces = 0;
  unsigned long long my_path = 0;
  uint32_t c_paths = 0, c_bounces = 0, c_dens = 0, c_alb = 0, c_esc = 0;

  for (;;) {
    // ------------------------------------------------------------ regeneration
    unsigned idle = __ballot_sync(FULL, state == S_IDLE);
    if (idle) {
      if (!exhausted) {
        int n = __popc(idle);
        int leader = __ffs(idle) - 1;
        unsigned long long base = 0;
        if ((int)lane == leader) base = atomicAdd(P.head, (unsigned long long)n);
        base = __shfl_sync(FULL, base, leader);
        if (state == S_IDLE) {
          unsigned long long g = base + __popc(idle & lane_lt);
          if (g < total) {
            // which tile / path / stream
            uint32_t j = 0;
            unsigned long long l = g;
            if (P.n_launch_tiles > 1) {
              j = (uint32_t)(g / per_tile);
              l = g - (unsigned long long)j * per_tile;
            }
            unsigned long long path_id = P.path_begin + l;
            uint32_t k = P.tile_first + j * P.tile_stride;
            uint32_t off_x = P.off_x, off_y = P.off_y;
            if (P.tile_origins) {
              uint2 og = P.tile_origins[k];
              off_x = og.x, off_y = og.y;
            }
            uint32_t seed_k = P.seed + k * P.seed_step;
            if (RNGM == RNG_XORWOW_PATH) rng.init((int32_t)(seed_k + (uint32_t)path_id));
            if (RNGM == RNG_PHILOX) rng.init((unsigned long long)seed_k + path_id);
            uint32_t image_id = (uint32_t)(path_id % P.npix);
            float u0 = rng.next();
            float u1 = rng.next();
            camera_ray(P.cam, image_id, off_x, off_y, u0, u1, o, d);
            thr_x = thr_y = thr_z = 1.f;
            bounces = 0;
            my_path = g;
            if (P.out_full) {
              uint32_t px = image_id % P.tile_w, py = image_id / P.tile_w;
              out_idx = (py + off_y) * P.out_stride + (px + off_x);
            } else {
              out_idx = image_id;
            }
            if (COUNT) ++c_paths;
            state = S_ISECT;
          } else {
            state = S_DONE;
          }
        }
        exhausted = (base + (unsigned long long)n >= total);
      } else if (state == S_IDLE) {
        state = S_DONE;
      }
    }
    if (__all_sync(FULL, state == S_DONE)) break;

    // ------------------------------------------------------------ intersect (A5)
    if (state == S_ISECT) {
      if (COUNT) ++c_bounces;
      V3 normal = v3(0, 0, 0);
      bool inside;
      if (!box_intersect(P.med.box_min, P.med.box_max, o, d, dist, normal, inside)) {
        // escaped: throughput * Le, Le == 1 (Medium.h:174-177); A13
        float rx = thr_x * 1.f, ry = thr_y * 1.f, rz = thr_z * 1.f;
        if (P.per_path) P.per_path[my_path] = make_float4(rx, ry, rz, 1.f);
        if (P.out) {
          float4* px = P.out + out_idx;
          atomicAdd(&px->x, rx);
          atomicAdd(&px->y, ry);
          atomicAdd(&px->z, rz);
          px->w = 1.f;
        }
        if (COUNT) ++c_esc;
        if (P.rr_after_escape && P.rr) (void)rng.next();
        state = S_IDLE;
      } else if (!inside) {
        ncode = code_from_normal(normal);
        state = S_BOUNDARY;
      } else {
        ncode = code_from_normal(normal);
        t = 0.f;
        state = S_TRACK;
      }
    }

    // ------------------------------------------------------------ Woodcock steps (A6, A7)
    {
      const int n_parked = __popc(__ballot_sync(FULL, state == S_DONE));
      // at least one step per outer iteration, then yield to the event phase as soon
      // as `loop_threshold` lanes are waiting for it
      for (bool first = true;; first = false) {
        unsigned trk = __ballot_sync(FULL, state == S_TRACK);
        if (trk == 0) break;
        int waiting = 32 - n_parked - __popc(trk);
        if (!first && waiting >= P.loop_threshold) break;
        if (state == S_TRACK) {
          float u = rng.next();
          t += -logf(fmaxf(u, CVR_EPS)) * inv_max_sigmat;
          V3 coord = (o + (t * d)) - q;
          float event_density = P.med.scale * density_lookup<LAYOUT>(P.med, coord);
          if (COUNT) ++c_dens;
          bool go_on = (t <= dist);
          if (go_on) go_on = (event_density * inv_max_sigmat < rng.next());
          if (!go_on) state = (t < dist) ? S_SCATTER : S_BOUNDARY;
        }
      }
    }

    // ------------------------------------------------------------ events
    if (state == S_SCATTER || state == S_BOUNDARY) {
      if (state == S_SCATTER) {
        // NaiveVolPTsk_kernel.cuh:67-71 / RegenerationVolPTsk_kernel.cuh:212-216
        if (P.pullback)
          o = o + d * t - d * CVR_EPS;
        else
          o = o + d * t;
        V3 ac = (o - P.med.box_min) / (P.med.box_max - P.med.box_min);
        V3 albedo = albedo_lookup<LAYOUT>(P.med, ac);
        if (COUNT) ++c_alb;
        thr_x = thr_x * albedo.x, thr_y = thr_y * albedo.y, thr_z = thr_z * albedo.z;
        float e1 = rng.next();
        float e2 = rng.next();
        d = hg_sample(d, P.med.hg_g, e1, e2);
      } else {
        // NaiveVolPTsk_kernel.cuh:50-65
        Frame frame;
        frame.from_z(normal_from_code(ncode));
        V3 dir = frame.to_local(normalize(v3(-d.x, -d.y, -d.z)));
        o = o + d * dist;
        float weight = 1;
        if (ggx_sample(P.med.alpha_x, P.med.alpha_y, P.med.eta, dir, rng, d, weight)) {
          thr_x *= weight, thr_y *= weight, thr_z *= weight;
          d = frame.to_world(d);
          o = o + d * CVR_EPS;
        }
      }
      // Russian roulette (NaiveVolPTsk_kernel.cuh:75-84)
      state = S_ISECT;
      if (P.rr) {
        float p_survive = fminf(1.f, fmaxf(fmaxf(thr_x, thr_y), thr_z));
        if (rng.next() > p_survive) state = S_IDLE;
        thr_x = thr_x * 1.f / p_survive;
        thr_y = thr_y * 1.f / p_survive;
        thr_z = thr_z * 1.f / p_survive;
      }
      ++bounces;
      if (P.max_bounces && bounces >= P.max_bounces) state = S_IDLE;
    }
  }

  if (COUNT) {
    // one atomic per warp per counter
    for (int s = 16; s > 0; s >>= 1) {
      c_paths += __shfl_xor_sync(FULL, c_paths, s);
      c_bounces += __shfl_xor_sync(FULL, c_bounces, s);
      c_dens += __shfl_xor_sync(FULL, c_dens, s);
      c_alb += __shfl_xor_sync(FULL, c_alb, s);
      c_esc += __shfl_xor_sync(FULL, c_esc, s);
    }
    if (lane == 0) {
      atomicAdd(&P.ctr->paths, (unsigned long long)c_paths);
      atomicAdd(&P.ctr->bounces, (unsigned long long)c_bounces);
      atomicAdd(&P.ctr->density_lookups, (unsigned long long)c_dens);
      atomicAdd(&P.ctr->albedo_lookups, (unsigned long long)c_alb);
      atomicAdd(&P.ctr->escaped, (unsigned long long)c_esc);
    }
  }
}

// ---------------------------------------------------------------- layout builders
// cell (kx,ky,kz) <- the 8 values the reference's 8 texture fetches return for
// x1 = kx-1 (see cell_index in cvr_device.cuh).
CVR_DEV int cell_lo(int k, int n) { return k == 0 ? n - 1 : k - 1; }
CVR_DEV int cell_hi(int k, int n) { return k < n - 1 ? k : n - 1; }

__global__ void k_build_density_cells(const float* __restrict__ D, int nx, int ny, int nz,
                                      float4* __restrict__ cells) {
  size_t ncell = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < ncell;
       i += (size_t)gridDim.x * blockDim.x) {
    int kx = (int)(i % (size_t)(nx + 1));
    size_t r = i / (size_t)(nx + 1);
    int ky = (int)(r % (size_t)(ny + 1));
    int kz = (int)(r / (size_t)(ny + 1));
    size_t X1 = cell_lo(kx, nx), X2 = cell_hi(kx, nx);
    size_t Y1 = cell_lo(ky, ny), Y2 = cell_hi(ky, ny);
    size_t Z1 = cell_lo(kz, nz), Z2 = cell_hi(kz, nz);
    size_t sx = (size_t)nx, sxy = (size_t)nx * ny;
    float4 a, b;
    a.x = D[X1 + sx * Y1 + sxy * Z1], a.y = D[X2 + sx * Y1 + sxy * Z1];
    a.z = D[X1 + sx * Y2 + sxy * Z1], a.w = D[X2 + sx * Y2 + sxy * Z1];
    b.x = D[X1 + sx * Y1 + sxy * Z2], b.y = D[X2 + sx * Y1 + sxy * Z2];
    b.z = D[X1 + sx * Y2 + sxy * Z2], b.w = D[X2 + sx * Y2 + sxy * Z2];
    cells[2 * i] = a;
    cells[2 * i + 1] = b;
  }
}

__global__ void k_build_albedo_cells(const float4* __restrict__ A, int nx, int ny, int nz,
                                     float4* __restrict__ cells) {
  size_t ncell = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
  // 8 threads per cell: one corner each -> 128-byte coalesced stores
  size_t nthr = ncell * 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nthr;
       i += (size_t)gridDim.x * blockDim.x) {
    int corner = (int)(i & 7);
    size_t c = i >> 3;
    int kx = (int)(c % (size_t)(nx + 1));
    size_t r = c / (size_t)(nx + 1);
    int ky = (int)(r % (size_t)(ny + 1));
    int kz = (int)(r / (size_t)(ny + 1));
    size_t X = (corner & 1) ? cell_hi(kx, nx) : cell_lo(kx, nx);
    size_t Y = (corner & 2) ? cell_hi(ky, ny) : cell_lo(ky, ny);
    size_t Z = (corner & 4) ? cell_hi(kz, nz) : cell_lo(kz, nz);
    cells[i] = A[X + (size_t)nx * (Y + (size_t)ny * Z)];
  }
}

// ---------------------------------------------------------------- resolve (A16)
// ImageBufferTransfer.cu:6-18 with UtilityFunctors::Scale (Utilities.h:6-15): every
// float of the tile divided by `scale`, written at the tile origin of the image.
__global__ void k_resolve_tile(const float4* __restrict__ tile, uint32_t tile_w, uint32_t tile_h,
                               uint32_t in_stride, uint32_t in_off_x, uint32_t in_off_y,
                               float4* __restrict__ image, uint32_t full_w, uint32_t off_x,
                               uint32_t off_y, float scale) {
  uint32_t n = tile_w * tile_h;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t x = i % tile_w, y = i / tile_w;
    float4 v = tile[(size_t)(y + in_off_y) * in_stride + (x + in_off_x)];
    v.x = v.x / scale, v.y = v.y / scale, v.z = v.z / scale, v.w = v.w / scale;
    image[(size_t)(y + off_y) * full_w + (x + off_x)] = v;
  }
}

// ---------------------------------------------------------------- debug kernels
__global__ void k_rng_kat(const int32_t* seeds, int n_seeds, int n, uint32_t* words, float* uni) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  Xorwow a, b;
  a.init(seeds[i]);
  b.init(seeds[i]);
  for (int k = 0; k < n; ++k) {
    words[(size_t)i * n + k] = a.next_u32();
    uni[(size_t)i * n + k] = b.next();
  }
}

template <int LAYOUT>
__global__ void k_debug_lookup(MediumParams m, const float* p, int n, float* dens, float* alb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3 c = v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  dens[i] = density_lookup<LAYOUT>(m, c);
  V3 a = albedo_lookup<LAYOUT>(m, c);
  alb[3 * i] = a.x, alb[3 * i + 1] = a.y, alb[3 * i + 2] = a.z;
}

}  // namespace cvr
