// cvr_volume.cuh -- device-side builders of the density lookup layouts from a voxel ACCESSOR
// (a functor voxel(x,y,z) -> value): procedural volumes generated in place (1024^3 dense,
// 2048^3 sparse -- too large to stage through host memory) and VDB-style leaves re-laid into
// bricks without densifying.  The accessor answers the same question the reference's
// texture fetch does (value of voxel (x,y,z), Volume.h:51-58), so the cells built here hold
// exactly the corner values -- including the wrap/clamp quirk Q2 -- of cell8 cells built from
// a dense array (k_build_density_cells).
#pragma once
#include "cvr_kernels.cuh"
#include "cvr_noise.h"

namespace cvr {

// ---------------------------------------------------------------- accessors
struct DenseAccessor {  // a dense x-fastest device array
  const float* D;
  int nx, ny, nz;
  __device__ float operator()(int x, int y, int z) const { return D[x + (size_t)nx * (y + (size_t)ny * z)]; }
  __device__ bool brick_active(int, int, int) const { return true; }
  __device__ float eval(int x, int y, int z, bool) const { return (*this)(x, y, z); }
};

struct FbmAccessor {  // "fbm" / "sparsefbm" (cvr_noise.h)
  uint32_t seed;
  int sparse;
  __device__ float operator()(int x, int y, int z) const {
    return sparse ? cvrnoise::sparsefbm_density(x, y, z, seed) : cvrnoise::fbm_density(x, y, z, seed);
  }
  // may the 8^3 voxel brick (bx,by,bz) hold a non-zero value?
  __device__ bool brick_active(int bx, int by, int bz) const {
    return sparse ? cvrnoise::sparse_brick_active(bx, by, bz, seed) : true;
  }
  // value of a voxel whose brick_active() answer is already known (the builders evaluate the
  // brick mask once per voxel brick, not once per voxel)
  __device__ float eval(int x, int y, int z, bool active) const {
    if (!sparse) return cvrnoise::fbm_density(x, y, z, seed);
    return active ? fmaxf(cvrnoise::fbm_density(x, y, z, seed), 1.0f / 64.0f) : 0.f;
  }
};

// VDB-style leaves: 8^3 voxels each, value n = (x&7)<<6 | (y&7)<<3 | (z&7) (the OpenVDB leaf
// order), origins in absolute index space.  Voxel (0,0,0) of the volume is index-space
// coordinate `base + off` where base is a multiple of 8; table covers the leaf grid from base.
struct LeafAccessor {
  const uint32_t* table;  // leaf number + 1, or 0
  const float* values;    // 512 per leaf, inactive voxels already resolved (0)
  int lbx, lby, lbz;      // leaf-grid dims
  int offx, offy, offz;   // 0..7
  __device__ float operator()(int x, int y, int z) const {
    x += offx, y += offy, z += offz;
    const uint32_t li = table[(x >> 3) + (size_t)lbx * ((y >> 3) + (size_t)lby * (z >> 3))];
    if (!li) return 0.f;
    return values[(size_t)(li - 1) * 512 + (((x & 7) << 6) | ((y & 7) << 3) | (z & 7))];
  }
  // here (bx,by,bz) is a brick of the LEAF grid
  __device__ bool leaf_present(int bx, int by, int bz) const {
    if (bx < 0 || by < 0 || bz < 0 || bx >= lbx || by >= lby || bz >= lbz) return false;
    return table[bx + (size_t)lbx * (by + (size_t)lby * bz)] != 0;
  }
};

// the 8 corner values of cell (kx,ky,kz) in the cell8 corner order (cvr_device.cuh)
template <class Acc>
__device__ __forceinline__ void cell_corners(const Acc& acc, int nx, int ny, int nz, int kx, int ky, int kz, float4& a,
                                             float4& b) {
  const int X1 = cell_lo(kx, nx), X2 = cell_hi(kx, nx);
  const int Y1 = cell_lo(ky, ny), Y2 = cell_hi(ky, ny);
  const int Z1 = cell_lo(kz, nz), Z2 = cell_hi(kz, nz);
  a.x = acc(X1, Y1, Z1), a.y = acc(X1, Y1, Z2), a.z = acc(X2, Y1, Z1), a.w = acc(X2, Y1, Z2);
  b.x = acc(X1, Y2, Z1), b.y = acc(X1, Y2, Z2), b.z = acc(X2, Y2, Z1), b.w = acc(X2, Y2, Z2);
}

// ---------------------------------------------------------------- dense cell8 from an accessor
template <class Acc>
__global__ void k_build_cells_fn(Acc acc, int nx, int ny, int nz, float4* __restrict__ cells) {
  const size_t ncell = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < ncell; i += (size_t)gridDim.x * blockDim.x) {
    const int kx = (int)(i % (size_t)(nx + 1));
    const size_t r = i / (size_t)(nx + 1);
    const int ky = (int)(r % (size_t)(ny + 1)), kz = (int)(r / (size_t)(ny + 1));
    float4 a, b;
    cell_corners(acc, nx, ny, nz, kx, ky, kz, a, b);
    cells[2 * i] = a, cells[2 * i + 1] = b;
  }
}

// ---------------------------------------------------------------- sparse bricks
// Voxel coordinates a cell brick b touches along one axis: cells k = 8b .. min(8b+7, n) use
// voxels x1 = k-1 (k = 0 wraps to n-1, Q2) and x2 = min(k, n-1).  Returns the inclusive
// voxel range [lo, hi] and whether voxel n-1 is touched through the wrap.
__device__ __forceinline__ void brick_voxel_range(int b, int n, int& lo, int& hi, bool& wrap) {
  const int k0 = 8 * b, k1 = min(8 * b + 7, n);
  wrap = (k0 == 0);
  lo = max(k0 - 1, 0);
  hi = min(k1, n - 1);
}

// Pass 1: a brick gets a slot iff a voxel brick it touches may be non-zero.  (Conservative:
// the slot order depends on the atomic order; the lookup results do not.)
template <class Acc>
__global__ void k_brick_slots_fn(Acc acc, int nx, int ny, int nz, uint32_t bmx, uint32_t bmy, uint32_t bmz,
                                 uint32_t* __restrict__ table, uint32_t* counter) {
  const size_t nb = (size_t)bmx * bmy * bmz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nb; i += (size_t)gridDim.x * blockDim.x) {
    const int bx = (int)(i % bmx), by = (int)((i / bmx) % bmy), bz = (int)(i / ((size_t)bmx * bmy));
    int lo[3], hi[3];
    bool wrap[3];
    brick_voxel_range(bx, nx, lo[0], hi[0], wrap[0]);
    brick_voxel_range(by, ny, lo[1], hi[1], wrap[1]);
    brick_voxel_range(bz, nz, lo[2], hi[2], wrap[2]);
    const int n[3] = {nx, ny, nz};
    // candidate voxel-brick coordinates per axis: the range, plus the far-edge brick on a wrap
    int cand[3][4], nc[3];
    for (int a = 0; a < 3; ++a) {
      nc[a] = 0;
      for (int v = lo[a] >> 3; v <= (hi[a] >> 3); ++v) cand[a][nc[a]++] = v;
      if (wrap[a] && ((n[a] - 1) >> 3) > (hi[a] >> 3)) cand[a][nc[a]++] = (n[a] - 1) >> 3;
    }
    bool any = false;
    for (int ix = 0; ix < nc[0] && !any; ++ix)
      for (int iy = 0; iy < nc[1] && !any; ++iy)
        for (int iz = 0; iz < nc[2] && !any; ++iz) any = acc.brick_active(cand[0][ix], cand[1][iy], cand[2][iz]);
    table[i] = any ? atomicAdd(counter, 1u) + 1u : 0u;
  }
}

// Pass 2: one CTA of 512 threads per brick-grid entry; entries without a slot return at once.
// The 512 cells of a brick share a 9^3 window of voxels (j-th coordinate of the window along an
// axis = voxel 8b-1+j, wrapped to n-1 below 0 and clamped to n-1 above: exactly cell_lo / cell_hi),
// so every voxel is evaluated once into shared memory -- and the brick mask of a sparse
// procedural volume once per voxel brick -- instead of 8 times per cell.
__device__ __forceinline__ int window_voxel(int b, int j, int n) {
  const int v = 8 * b - 1 + j;
  return v < 0 ? n - 1 : (v > n - 1 ? n - 1 : v);
}
template <class Acc>
__global__ void __launch_bounds__(512) k_build_bricks_fn(Acc acc, int nx, int ny, int nz, uint32_t bmx, uint32_t bmy,
                                                         const uint32_t* __restrict__ table,
                                                         float4* __restrict__ bricks, uint32_t max_slots) {
  const size_t i = blockIdx.x + (size_t)blockIdx.y * gridDim.x;
  const uint32_t slot = table[i];
  if (slot == 0u || slot > max_slots) return;
  const int bx = (int)(i % bmx), by = (int)((i / bmx) % bmy), bz = (int)(i / ((size_t)bmx * bmy));
  __shared__ float s_v[9 * 9 * 9];
  __shared__ int s_brick[3][9];     // voxel-brick coordinate of window entry j, per axis
  __shared__ unsigned char s_mask[27];  // brick mask of every (distinct brick per axis) combination
  __shared__ int s_rank[3][9], s_nd[3], s_dist[3][3];
  const unsigned t = threadIdx.x;
  if (t < 27) {
    const int a = t / 9, j = t % 9;
    const int n = a == 0 ? nx : a == 1 ? ny : nz, b = a == 0 ? bx : a == 1 ? by : bz;
    s_brick[a][j] = window_voxel(b, j, n) >> 3;
  }
  __syncthreads();
  if (t < 3) {  // distinct voxel bricks along axis t (at most 3: below, this, far edge on a wrap)
    int nd = 0;
    for (int j = 0; j < 9; ++j) {
      int r = -1;
      for (int k = 0; k < nd; ++k)
        if (s_dist[t][k] == s_brick[t][j]) r = k;
      if (r < 0) r = nd, s_dist[t][nd++] = s_brick[t][j];
      s_rank[t][j] = r;
    }
    s_nd[t] = nd;
  }
  __syncthreads();
  if (t < 27) {
    const int rx = t % 3, ry = (t / 3) % 3, rz = t / 9;
    s_mask[t] = (rx < s_nd[0] && ry < s_nd[1] && rz < s_nd[2]) ? acc.brick_active(s_dist[0][rx], s_dist[1][ry], s_dist[2][rz]) : 0;
  }
  __syncthreads();
  for (unsigned v = t; v < 729u; v += 512u) {
    const int jx = v % 9, jy = (v / 9) % 9, jz = v / 81;
    const bool active = s_mask[s_rank[0][jx] + 3 * s_rank[1][jy] + 9 * s_rank[2][jz]] != 0;
    s_v[v] = acc.eval(window_voxel(bx, jx, nx), window_voxel(by, jy, ny), window_voxel(bz, jz, nz), active);
  }
  __syncthreads();
  const int tx = (int)(t & 7u), ty = (int)((t >> 3) & 7u), tz = (int)(t >> 6);
  float4 a = make_float4(0, 0, 0, 0), b = a;
  if (8 * bx + tx <= nx && 8 * by + ty <= ny && 8 * bz + tz <= nz) {
    // corner (x1|x2, y1|y2, z1|z2) = window entry (t | t+1) per axis; cell8 corner order
    auto V = [&](int jx, int jy, int jz) { return s_v[jx + 9 * (jy + 9 * jz)]; };
    a.x = V(tx, ty, tz), a.y = V(tx, ty, tz + 1), a.z = V(tx + 1, ty, tz), a.w = V(tx + 1, ty, tz + 1);
    b.x = V(tx, ty + 1, tz), b.y = V(tx, ty + 1, tz + 1), b.z = V(tx + 1, ty + 1, tz), b.w = V(tx + 1, ty + 1, tz + 1);
  }
  float4* out = bricks + ((size_t)slot * 512u + t) * 2u;
  out[0] = a, out[1] = b;
}

// majorant per brick-grid entry (tracking=local) and the global maximum (max_density)
__global__ void k_brick_majorant(const uint32_t* __restrict__ table, const float4* __restrict__ bricks, size_t nb,
                                 float* __restrict__ maj, unsigned int* global_max_bits) {
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const unsigned lane = threadIdx.x & 31u;
  if (warp >= nb) return;
  const uint32_t slot = table[warp];
  float m = 0.f;
  if (slot) {
    const float4* c = bricks + (size_t)slot * 1024u;
    for (unsigned i = lane; i < 1024u; i += 32u) {
      const float4 v = c[i];
      m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
  }
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if (lane == 0) {
    maj[warp] = m;
    if (m > 0.f) atomicMax(global_max_bits, __float_as_uint(m));  // non-negative floats order like their bits
  }
}

// leaf number + 1 into the leaf-grid table
__global__ void k_fill_leaf_table(const int32_t* __restrict__ origins, size_t n_leaves, int basex, int basey, int basez,
                                  int lbx, int lby, int lbz, uint32_t* __restrict__ table) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n_leaves) return;
  const int bx = (origins[3 * i] - basex) >> 3, by = (origins[3 * i + 1] - basey) >> 3, bz = (origins[3 * i + 2] - basez) >> 3;
  if (bx < 0 || by < 0 || bz < 0 || bx >= lbx || by >= lby || bz >= lbz) return;
  table[bx + (size_t)lbx * (by + (size_t)lby * bz)] = (uint32_t)i + 1u;
}

// Accessor over leaves whose brick_active() speaks the VOXEL-brick grid of the volume
// (coordinates relative to voxel (0,0,0)): a relative brick can straddle up to 8 leaves when
// the volume origin is not a multiple of 8.
struct LeafVolumeAccessor {
  LeafAccessor L;
  __device__ float operator()(int x, int y, int z) const { return L(x, y, z); }
  __device__ float eval(int x, int y, int z, bool) const { return L(x, y, z); }
  __device__ bool brick_active(int bx, int by, int bz) const {
    const int x0 = 8 * bx + L.offx, y0 = 8 * by + L.offy, z0 = 8 * bz + L.offz;
    for (int dx = 0; dx < 2; ++dx)
      for (int dy = 0; dy < 2; ++dy)
        for (int dz = 0; dz < 2; ++dz)
          if (L.leaf_present((x0 + 7 * dx) >> 3, (y0 + 7 * dy) >> 3, (z0 + 7 * dz) >> 3)) return true;
    return false;
  }
};

}  // namespace cvr
