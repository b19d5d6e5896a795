// cvr_noise.h -- the procedural density functions of the synthetic scenes, shared by the host
// generator (cvr_synth.cpp) and the device-side layout builders (cvr_volume.cuh) so that a
// volume generated on the GPU (1024^3 dense, 2048^3 sparse: too large to stage through host
// memory) has exactly the voxels the host generator produces for the same parameters.
// Every multiply/add is explicit (no FMA contraction on the device) for that reason.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define CVR_HD __host__ __device__ __forceinline__
#else
#define CVR_HD inline
#endif

namespace cvrnoise {

#ifdef __CUDA_ARCH__
CVR_HD float mul(float a, float b) { return __fmul_rn(a, b); }
CVR_HD float add(float a, float b) { return __fadd_rn(a, b); }
CVR_HD float sub(float a, float b) { return __fsub_rn(a, b); }
#else
CVR_HD float mul(float a, float b) { return a * b; }
CVR_HD float add(float a, float b) { return a + b; }
CVR_HD float sub(float a, float b) { return a - b; }
#endif

CVR_HD uint32_t hash32(uint32_t a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

CVR_HD float lattice(int x, int y, int z, uint32_t octave, uint32_t seed) {
  uint32_t key = ((uint32_t)x & 1023u) + 1024u * (((uint32_t)y & 1023u) + 1024u * ((uint32_t)z & 1023u));
  uint32_t hv = hash32(key ^ (octave * 0x9e3779b9u) ^ seed);
  return (float)(hv >> 8) * (1.0f / 16777216.0f);
}

CVR_HD float fade(float t) { return mul(mul(t, t), sub(3.0f, mul(2.0f, t))); }
CVR_HD float lerp(float a, float b, float t) { return add(a, mul(t, sub(b, a))); }

// value noise at lattice-space position p
CVR_HD float vnoise(float px, float py, float pz, uint32_t octave, uint32_t seed) {
  float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
  int x = (int)fx, y = (int)fy, z = (int)fz;
  float u = fade(sub(px, fx)), v = fade(sub(py, fy)), w = fade(sub(pz, fz));
  float c000 = lattice(x, y, z, octave, seed), c100 = lattice(x + 1, y, z, octave, seed);
  float c010 = lattice(x, y + 1, z, octave, seed), c110 = lattice(x + 1, y + 1, z, octave, seed);
  float c001 = lattice(x, y, z + 1, octave, seed), c101 = lattice(x + 1, y, z + 1, octave, seed);
  float c011 = lattice(x, y + 1, z + 1, octave, seed), c111 = lattice(x + 1, y + 1, z + 1, octave, seed);
  float a = lerp(c000, c100, u), b = lerp(c010, c110, u);
  float c = lerp(c001, c101, u), d = lerp(c011, c111, u);
  return lerp(lerp(a, b, v), lerp(c, d, v), w);
}

// 5 octaves, lacunarity 2, gain 0.5, normalised to [0,1]
CVR_HD float fbm5(float px, float py, float pz, uint32_t seed) {
  float sum = 0.f, amp = 0.5f, norm = 0.f;
  for (uint32_t o = 0; o < 5; ++o) {
    sum = add(sum, mul(amp, vnoise(px, py, pz, o, seed)));
    norm = add(norm, amp);
    px = mul(px, 2.f), py = mul(py, 2.f), pz = mul(pz, 2.f);
    amp = mul(amp, 0.5f);
  }
  return sum / norm;
}

// "fbm": value-noise fBm with a 128-voxel base period, rho = max(0, fbm - 0.4) / 0.6
// 
CVR_HD float fbm_density(int x, int y, int z, uint32_t seed) {
  const float inv_period = 1.0f / 128.0f;
  float f = fbm5(mul((float)x, inv_period), mul((float)y, inv_period), mul((float)z, inv_period), seed);
  return fmaxf(0.f, sub(f, 0.4f)) / 0.6f;
}

// "sparsefbm": VDB-style sparse volume -- 8^3 voxel bricks are active where
// a coarse fBm over the brick grid (16-brick base period) exceeds a threshold tuned to ~3 %
// occupancy; inside active bricks the density is the fbm density, elsewhere exactly 0.
#define CVR_SPARSE_THRESHOLD 0.70f
CVR_HD bool sparse_brick_active(int bx, int by, int bz, uint32_t seed) {
  const float inv = 1.0f / 16.0f;
  return fbm5(mul((float)bx, inv), mul((float)by, inv), mul((float)bz, inv), seed ^ 0xb41c0de5u) > CVR_SPARSE_THRESHOLD;
}
CVR_HD float sparsefbm_density(int x, int y, int z, uint32_t seed) {
  if (!sparse_brick_active(x >> 3, y >> 3, z >> 3, seed)) return 0.f;
  // a floor keeps every voxel of an active brick non-zero, like a VDB leaf of active values
  return fmaxf(fbm_density(x, y, z, seed), 1.0f / 64.0f);
}

}  // namespace cvrnoise
