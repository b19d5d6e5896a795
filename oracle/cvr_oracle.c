/*
 * cvr_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see cvr_oracle.h).
 *
 * Plain C99 + pthreads, fp32 throughout, compiled with -ffp-contract=off so the
 * result does not depend on the host's FMA availability.  Every function cites
 * the reference lines (relative to /root/reference/implementation/src/) whose
 * behaviour it restates.  Nothing here is shared with the CUDA product path.
 */
#include "cvr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define CVRO_EPS 0.00001f               /* Defines.h:64 */
#define CVRO_PI 3.14159265358979323846f /* Defines.h:62 (rounds to the same float) */
#define CVRO_TWOPI 6.28318530717958647692f /* Defines.h:63 */

typedef struct {
  float x, y, z;
} v3;

/* ---- float3 helpers with the expression shapes of helper_math.h ---- */
static void log_decision(uint32_t kind, float a, float b); /* optional path log, defined below */
static void log_event(uint32_t code, uint32_t d);
static inline v3 V(float x, float y, float z) {
  v3 r = {x, y, z};
  return r;
}
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) {
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* helper_math.h:1055-1057: v * rsqrtf(dot(v,v)).  rsqrtf here is the CUDA
 * toolkit's HOST definition (crt/math_functions.hpp: (float)(1.0/sqrt((double)a))),
 * which is what the reference gets when host-compiled through nvcc; the device
 * intrinsic differs from it by at most 2 ulp. */
static inline float cvro_rsqrtf(float a) { return (float)(1.0 / sqrt((double)a)); }
static inline v3 vnormalize(v3 v) {
  float inv_len = cvro_rsqrtf(vdot(v, v));
  return vscale(v, inv_len);
}

/* ------------------------------------------------------------------------- */
/* RNG: curand_kernel.h _curand_init_scratch (seed only, subsequence = offset = 0) */
/* and curand(curandStateXORWOW_t*); Rng.h:22 passes an int seed, which C++     */
/* converts to unsigned long long by sign extension.                             */
/* ------------------------------------------------------------------------- */
void cvro_rng_init(cvro_rng* r, int32_t seed) {
  uint64_t s = (uint64_t)(int64_t)seed;
  uint32_t s0 = (uint32_t)s ^ 0xaad26b49u;
  uint32_t s1 = (uint32_t)(s >> 32) ^ 0xf7dcefddu;
  uint32_t t0 = 1099087573u * s0;
  uint32_t t1 = 2591861531u * s1;
  r->d = 6615241u + t1 + t0;
  r->v[0] = 123456789u + t0;
  r->v[1] = 362436069u ^ t0;
  r->v[2] = 521288629u + t1;
  r->v[3] = 88675123u ^ t1;
  r->v[4] = 5783321u + t0;
}

uint32_t cvro_rng_u32(cvro_rng* r) {
  uint32_t t = r->v[0] ^ (r->v[0] >> 2);
  r->v[0] = r->v[1];
  r->v[1] = r->v[2];
  r->v[2] = r->v[3];
  r->v[3] = r->v[4];
  r->v[4] = (r->v[4] ^ (r->v[4] << 4)) ^ (t ^ (t << 1));
  r->d += 362437u;
  return r->v[4] + r->d;
}

/* curand_uniform.h:69-72.  nvcc contracts x*c + c/2 into one fma on the device,
 * so the oracle states the fma explicitly (checked on the GPU by
 * tests/test_gpu_rng.py against cuRAND itself). */
float cvro_rng_float(cvro_rng* r) {
  uint32_t x = cvro_rng_u32(r);
  return fmaf((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
}

/* ------------------------------------------------------------------------- */
/* hashes (Utilities.cuh:157-171, Utilities.h:35-55)                             */
/* ------------------------------------------------------------------------- */
uint32_t cvro_utilhash(uint32_t a) {
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

static uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

uint32_t cvro_morton3d(float x, float y, float z) {
  x = fminf(fmaxf(x * 1024.0f, 0.0f), 1023.0f);
  y = fminf(fmaxf(y * 1024.0f, 0.0f), 1023.0f);
  z = fminf(fmaxf(z * 1024.0f, 0.0f), 1023.0f);
  return expand_bits10((uint32_t)x) * 4u + expand_bits10((uint32_t)y) * 2u +
         expand_bits10((uint32_t)z);
}

/* ------------------------------------------------------------------------- */
/* camera (Utilities.cuh:180-213, CVRMath.h:18-43)                               */
/* ------------------------------------------------------------------------- */
static uint32_t f2u_trunc(float f) {
  /* device cvt.rzi.u32.f32 saturates; NaN -> 0 */
  if (!(f > 0.0f)) return 0u;
  if (f >= 4294967296.0f) return 0xffffffffu;
  return (uint32_t)f;
}

void cvro_camera_ray(const cvro_camera* cam, uint32_t image_id, float u0, float u1,
                     float o[3], float d[3]) {
  /* NaiveVolPTsk_kernel.cuh:25-27 */
  float px = (float)(image_id % f2u_trunc(cam->resolution[0])) + (float)cam->offset[0];
  float py = floorf((float)image_id / cam->resolution[0]) + (float)cam->offset[1];
  /* indexToRaster: (pixel + rnd) * 2 / range - 1 */
  px = px + u0;
  py = py + u1;
  float rx = (px * 2.f / cam->pixel_index_range[0]) - 1.0f;
  float ry = (py * 2.f / cam->pixel_index_range[1]) - 1.0f;
  /* cameraGenerateRay */
  rx = cam->raster_to_view[0] * rx;
  ry = cam->raster_to_view[1] * ry;
  const float* m = cam->inv_view;
  /* mul(M, float4(0,0,0,1)): dot with each row */
  o[0] = 0.0f * m[0] + 0.0f * m[1] + 0.0f * m[2] + 1.0f * m[3];
  o[1] = 0.0f * m[4] + 0.0f * m[5] + 0.0f * m[6] + 1.0f * m[7];
  o[2] = 0.0f * m[8] + 0.0f * m[9] + 0.0f * m[10] + 1.0f * m[11];
  v3 dir = vnormalize(V(rx, ry, 1.0f));
  d[0] = vdot(dir, V(m[0], m[1], m[2]));
  d[1] = vdot(dir, V(m[4], m[5], m[6]));
  d[2] = vdot(dir, V(m[8], m[9], m[10]));
}

void cvro_default_camera(uint32_t res_x, uint32_t res_y, float fov_x,
                         float inv_view[12], float raster_to_view[2]) {
  /* Camera.h:25-37: columns right(1,0,0) up(0,-1,0) view(0,0,-1) pos(0,0,100);
   * CudaVolPath.cpp:71-84 transposes the 3x3 and moves the position to column 3. */
  static const float m[12] = {1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1, 100.0f};
  memcpy(inv_view, m, sizeof m);
  /* Camera.h:63-71 */
  float fov_y = ((float)res_y / (float)res_x) * fov_x;
  raster_to_view[0] = tanf(fov_x * CVRO_PI / 360.f);
  raster_to_view[1] = tanf(fov_y * CVRO_PI / 360.f);
}

/* Config.h:67-72 (integer division inside ceil => floor, Q6) and
 * CudaVolPath.cpp:15-19 */
void cvro_tile_table(uint32_t res_x, uint32_t res_y, uint32_t ntx, uint32_t nty,
                     uint32_t tile_dim[2], uint32_t* origins) {
  tile_dim[0] = (uint32_t)(int)ceil((double)(res_x / ntx));
  tile_dim[1] = (uint32_t)(int)ceil((double)(res_y / nty));
  for (uint32_t id = 0; id < ntx * nty; ++id) {
    origins[2 * id + 0] = tile_dim[0] * (id % ntx);
    origins[2 * id + 1] = tile_dim[1] * (uint32_t)(int)((float)id / (float)ntx);
  }
}

/* ------------------------------------------------------------------------- */
/* AABB slab test (Geometry.h:55-92).  `normal` and `inside` are in/out: when no */
/* equality matches the reference leaves the previous normal in place.           */
/* ------------------------------------------------------------------------- */
int cvro_aabb_intersect(const float bmin[3], const float bmax[3],
                        const float o[3], const float d[3], float* dist,
                        float normal[3], int* inside) {
  v3 invR = V(1.0f / d[0], 1.0f / d[1], 1.0f / d[2]);
  v3 tbot = vmul(invR, V(bmin[0] - o[0], bmin[1] - o[1], bmin[2] - o[2]));
  v3 ttop = vmul(invR, V(bmax[0] - o[0], bmax[1] - o[1], bmax[2] - o[2]));
  v3 tmin = V(fminf(ttop.x, tbot.x), fminf(ttop.y, tbot.y), fminf(ttop.z, tbot.z));
  v3 tmax = V(fmaxf(ttop.x, tbot.x), fmaxf(ttop.y, tbot.y), fmaxf(ttop.z, tbot.z));
  float largest_tmin = fmaxf(fmaxf(tmin.x, tmin.y), fmaxf(tmin.x, tmin.z));
  float smallest_tmax = fminf(fminf(tmax.x, tmax.y), fminf(tmax.x, tmax.z));
  float dd = (largest_tmin > CVRO_EPS) ? largest_tmin : smallest_tmax;
  *dist = dd;
  if (dd == ttop.x) {
    normal[0] = 1, normal[1] = 0, normal[2] = 0;
  } else if (dd == ttop.y) {
    normal[0] = 0, normal[1] = 1, normal[2] = 0;
  } else if (dd == ttop.z) {
    normal[0] = 0, normal[1] = 0, normal[2] = 1;
  } else if (dd == tbot.x) {
    normal[0] = -1, normal[1] = 0, normal[2] = 0;
  } else if (dd == tbot.y) {
    normal[0] = 0, normal[1] = -1, normal[2] = 0;
  } else if (dd == tbot.z) {
    normal[0] = 0, normal[1] = 0, normal[2] = -1;
  }
  *inside = (normal[0] * d[0] + normal[1] * d[1] + normal[2] * d[2]) > 0;
  return (smallest_tmax > largest_tmin) && (dd > 0);
}

/* ------------------------------------------------------------------------- */
/* trilinear lookups (Volume.h:40-69; texture = point filter, clamp,            */
/* unnormalised: CudaVolPath.cpp:168-181; get(uint,uint,uint):                  */
/* RenderKernelLauncher.cu:20-25, so a negative int wraps and clamps to the FAR  */
/* edge, Q2)                                                                    */
/* ------------------------------------------------------------------------- */
static int32_t f2i_sat(float f) {
  /* device cvt.rzi.s32.f32: saturating, NaN -> 0 */
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (int32_t)f;
}

static inline uint32_t clamp_texel(int32_t i, int32_t n) {
  uint32_t u = (uint32_t)i;
  return u > (uint32_t)(n - 1) ? (uint32_t)(n - 1) : u;
}

float cvro_density_lookup(const cvro_scene* sc, const float p[3]) {
  const int nx = sc->dnx, ny = sc->dny, nz = sc->dnz;
  /* volumeToGrid: p * (uint)(res - 1) */
  float cx = p[0] * (float)(uint32_t)(nx - 1);
  float cy = p[1] * (float)(uint32_t)(ny - 1);
  float cz = p[2] * (float)(uint32_t)(nz - 1);
  int32_t x1 = f2i_sat(floorf(cx)), y1 = f2i_sat(floorf(cy)), z1 = f2i_sat(floorf(cz));
  /* x1 + 1 on int: the reference would overflow at INT_MAX (UB); wrap like the GPU */
  int32_t x2 = (int32_t)((uint32_t)x1 + 1u), y2 = (int32_t)((uint32_t)y1 + 1u),
          z2 = (int32_t)((uint32_t)z1 + 1u);
  float fx = cx - (float)x1, fy = cy - (float)y1, fz = cz - (float)z1;
  float _fx = 1.0f - fx, _fy = 1.0f - fy, _fz = 1.0f - fz;
  size_t X1 = clamp_texel(x1, nx), X2 = clamp_texel(x2, nx);
  size_t Y1 = clamp_texel(y1, ny), Y2 = clamp_texel(y2, ny);
  size_t Z1 = clamp_texel(z1, nz), Z2 = clamp_texel(z2, nz);
  const float* D = sc->density;
#define AT(x, y, z) D[(x) + (size_t)nx * ((y) + (size_t)ny * (z))]
  float d000 = AT(X1, Y1, Z1), d001 = AT(X2, Y1, Z1), d010 = AT(X1, Y2, Z1),
        d011 = AT(X2, Y2, Z1), d100 = AT(X1, Y1, Z2), d101 = AT(X2, Y1, Z2),
        d110 = AT(X1, Y2, Z2), d111 = AT(X2, Y2, Z2);
#undef AT
  return ((d000 * _fx + d001 * fx) * _fy + (d010 * _fx + d011 * fx) * fy) * _fz +
         ((d100 * _fx + d101 * fx) * _fy + (d110 * _fx + d111 * fx) * fy) * fz;
}

void cvro_albedo_lookup(const cvro_scene* sc, const float p[3], float rgb[3]) {
  const int nx = sc->anx, ny = sc->any, nz = sc->anz;
  float cx = p[0] * (float)(uint32_t)(nx - 1);
  float cy = p[1] * (float)(uint32_t)(ny - 1);
  float cz = p[2] * (float)(uint32_t)(nz - 1);
  int32_t x1 = f2i_sat(floorf(cx)), y1 = f2i_sat(floorf(cy)), z1 = f2i_sat(floorf(cz));
  int32_t x2 = (int32_t)((uint32_t)x1 + 1u), y2 = (int32_t)((uint32_t)y1 + 1u),
          z2 = (int32_t)((uint32_t)z1 + 1u);
  float fx = cx - (float)x1, fy = cy - (float)y1, fz = cz - (float)z1;
  float _fx = 1.0f - fx, _fy = 1.0f - fy, _fz = 1.0f - fz;
  size_t X1 = clamp_texel(x1, nx), X2 = clamp_texel(x2, nx);
  size_t Y1 = clamp_texel(y1, ny), Y2 = clamp_texel(y2, ny);
  size_t Z1 = clamp_texel(z1, nz), Z2 = clamp_texel(z2, nz);
  const float* A = sc->albedo;
#define AT(x, y, z, c) A[4 * ((x) + (size_t)nx * ((y) + (size_t)ny * (z))) + (c)]
  for (int c = 0; c < 3; ++c) {
    float d000 = AT(X1, Y1, Z1, c), d001 = AT(X2, Y1, Z1, c), d010 = AT(X1, Y2, Z1, c),
          d011 = AT(X2, Y2, Z1, c), d100 = AT(X1, Y1, Z2, c), d101 = AT(X2, Y1, Z2, c),
          d110 = AT(X1, Y2, Z2, c), d111 = AT(X2, Y2, Z2, c);
    rgb[c] = ((d000 * _fx + d001 * fx) * _fy + (d010 * _fx + d011 * fx) * fy) * _fz +
             ((d100 * _fx + d101 * fx) * _fy + (d110 * _fx + d111 * fx) * fy) * fz;
  }
#undef AT
}

/* ------------------------------------------------------------------------- */
/* coordinate frame (CVRMath.h:69-75)                                            */
/* ------------------------------------------------------------------------- */
void cvro_frame_from_z(const float n[3], float x[3], float y[3], float z[3]) {
  v3 tz = vnormalize(V(n[0], n[1], n[2]));
  v3 tx = (fabsf(tz.x) > 0.99f) ? V(0, 1, 0) : V(1, 0, 0);
  v3 ty = vnormalize(vcross(tz, tx));
  v3 xx = vcross(ty, tz);
  x[0] = xx.x, x[1] = xx.y, x[2] = xx.z;
  y[0] = ty.x, y[1] = ty.y, y[2] = ty.z;
  z[0] = tz.x, z[1] = tz.y, z[2] = tz.z;
}

/* ------------------------------------------------------------------------- */
/* Henyey-Greenstein sampling (HG.h:11-24,46-63)                                 */
/* ------------------------------------------------------------------------- */
void cvro_hg_sample(const float dir[3], float g, float e1, float e2, float out[3]) {
  float cos_theta;
  if (fabsf(g) > CVRO_EPS) {
    float sqr = (1.0f - g * g) / (1.0f - g + 2.0f * g * e1);
    cos_theta = (1.0f + g * g - sqr * sqr) / (2.0f * fabsf(g));
  } else {
    cos_theta = 1.0f - 2.0f * e1;
  }
  float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
  float phi = CVRO_TWOPI * e2;
  v3 v = V(dir[0], dir[1], dir[2]);
  float inv_norm = 1.0f / sqrtf(v.x * v.x + v.z * v.z);
  v3 v1 = V(v.z * inv_norm, 0.0f, -v.x * inv_norm);
  v3 v2 = vcross(v, v1);
  float a = sin_theta * cosf(phi);
  float b = sin_theta * sinf(phi);
  v3 r = vadd(vadd(vscale(v1, a), vscale(v2, b)), vscale(v, cos_theta));
  out[0] = r.x, out[1] = r.y, out[2] = r.z;
}

/* ------------------------------------------------------------------------- */
/* rough dielectric boundary (GGX.h:13-38,40-50,85-181,213-255,265-326)          */
/* ------------------------------------------------------------------------- */
float cvro_fresnel_dielectric(float eta, float ndotwi, float* p_ndotwt) {
  if (eta == 1) {
    *p_ndotwt = -ndotwi;
    return 0.0f;
  }
  float scale = (ndotwi > 0) ? 1 / eta : eta;
  float sin_sqr = (1 - (ndotwi * ndotwi));
  float ndotwt_sqr = 1 - (sin_sqr * scale * scale);
  if (ndotwt_sqr <= 0.0f) {
    *p_ndotwt = 0.0f;
    return 1.0f;
  }
  float abs_ndotwi = fabsf(ndotwi);
  float abs_ndotwt = sqrtf(ndotwt_sqr);
  float Rs = (abs_ndotwi - eta * abs_ndotwt) / (abs_ndotwi + eta * abs_ndotwt);
  float Rp = (eta * abs_ndotwi - abs_ndotwt) / (eta * abs_ndotwi + abs_ndotwt);
  *p_ndotwt = (ndotwi > 0) ? -abs_ndotwt : abs_ndotwt;
  return 0.5f * (Rs * Rs + Rp * Rp);
}

static void sample_visible11(float theta_i, float sx_in, float sy_in, float* slope_x,
                             float* slope_y) {
  float phi = 2 * CVRO_PI * sy_in;
  if (theta_i < 1e-4f) {
    float r = sqrtf(fmaxf(0.0f, sx_in / (1 - sx_in)));
    float sin_phi = sinf(phi), cos_phi = cosf(phi);
    *slope_x = r * cos_phi;
    *slope_y = r * sin_phi;
    return;
  }
  float tan_theta_i = tanf(theta_i);
  float a = 1 / tan_theta_i;
  a = 1.0f + (1.0f / (a * a));
  float G1 = 2.0f / (1.0f + sqrtf(a));
  float A = (2.0f * sx_in / G1) - 1.0f;
  if (fabsf(A) == 1) A -= copysignf(1.0f, A) * CVRO_EPS;
  float tmp = 1.0f / (A * A - 1.0f);
  float B = tan_theta_i;
  float D = sqrtf(fmaxf(0.0f, (B * B * tmp * tmp) - ((A * A - B * B) * tmp)));
  float s1 = (B * tmp) - D;
  float s2 = (B * tmp) + D;
  float sx = (A < 0.0f || s2 > 1.0f / tan_theta_i) ? s1 : s2;
  float S, u = sy_in;
  if (u > 0.5f) {
    S = 1.0f;
    u = 2.0f * (u - 0.5f);
  } else {
    S = -1.0f;
    u = 2.0f * (0.5f - u);
  }
  float z = (u * (u * (u * (-(float)0.365728915865723) + (float)0.790235037209296) -
                  (float)0.424965825137544) +
             (float)0.000152998850436920) /
            (u * (u * (u * (u * (float)0.169507819808272 - (float)0.397203533833404) -
                       (float)0.232500544458471) +
                  (float)1) -
             (float)0.539825872510702);
  *slope_x = sx;
  *slope_y = S * z * sqrtf(1.0f + (sx * sx));
}

static v3 ggx_sample_vndf(v3 wi_in, const float alpha[2], float u1, float u2) {
  v3 wi = vnormalize(V(alpha[0] * wi_in.x, alpha[1] * wi_in.y, wi_in.z));
  float theta = 0, phi = 0;
  if (wi.z < (float)0.999999) {
    theta = acosf(wi.z);
    phi = atan2f(wi.y, wi.x);
  }
  float sin_phi = sinf(phi), cos_phi = cosf(phi);
  float sx, sy;
  sample_visible11(theta, u1, u2, &sx, &sy);
  float rx = (cos_phi * sx) - (sin_phi * sy);
  float ry = (sin_phi * sx) + (cos_phi * sy);
  rx *= alpha[0];
  ry *= alpha[1];
  float nrm = 1.f / sqrtf((rx * rx) + (ry * ry) + 1.0f);
  return V(-rx * nrm, -ry * nrm, nrm);
}

float cvro_ggx_g1(const float alpha[2], const float vv[3], const float mm[3]) {
  v3 v = V(vv[0], vv[1], vv[2]), m = V(mm[0], mm[1], mm[2]);
  if (vdot(v, m) * v.z <= 0) return 0.0f;
  float temp = 1 - (v.z * v.z);
  if (temp <= 0.0f) return 0.0f;
  float tn = sqrtf(temp) / v.z;
  tn = fabsf(tn);
  if (tn == 0.0f) return 1.0f;
  /* projectRoughness (GGX.h:213-225) */
  float proj;
  float inv_sin2 = 1 / (1.0f - v.z * v.z);
  if (alpha[0] == alpha[1] || inv_sin2 <= 0) {
    proj = alpha[0];
  } else {
    float cos_phi2 = v.x * v.x * inv_sin2;
    float sin_phi2 = v.y * v.y * inv_sin2;
    proj = sqrtf((cos_phi2 * alpha[0] * alpha[0]) + (sin_phi2 * alpha[1] * alpha[1]));
  }
  float root = proj * tn;
  return 2.0f / (1.0f + sqrtf(1.0f + (root * root)));
}

int cvro_ggx_sample(const float alpha[2], float eta, const float wi_in[3],
                    const float u[3], float wo[3], float* weight, int* n_used) {
  v3 wi = V(wi_in[0], wi_in[1], wi_in[2]);
  *n_used = 0;
  if (wi.z == 0.f) {
    *weight = 0;
    return 0;
  }
  *weight = 1.0f;
  float sign = wi.z / fabsf(wi.z);
  v3 wh = ggx_sample_vndf(vscale(wi, sign), alpha, u[0], u[1]);
  *n_used = 2;
  float whdotwt = NAN;
  float whdotwi = vdot(wh, wi);
  float F = cvro_fresnel_dielectric(eta, whdotwi, &whdotwt);
  *n_used = 3;
  log_decision(CVRO_DEC_FRESNEL, u[2], F);
  if (u[2] <= F) {
    /* reflect (GGX.h:40-43): 2*c*wh - wi, written through *wo before the check */
    v3 r = vsub(vscale(wh, 2.f * whdotwi), wi);
    wo[0] = r.x, wo[1] = r.y, wo[2] = r.z;
    if (wi.z * r.z <= 0) {
      *weight = 0.0f;
      return 0;
    }
  } else {
    if (whdotwt == 0.0f) {
      *weight = 0.0f;
      return 0;
    }
    /* refract (GGX.h:45-50) */
    float e = eta;
    if (whdotwt < 0) e = 1 / e;
    v3 r = vsub(vscale(wh, whdotwi * e + whdotwt), vscale(wi, e));
    wo[0] = r.x, wo[1] = r.y, wo[2] = r.z;
    if (wi.z * r.z >= 0) {
      *weight = 0.0f;
      return 0;
    }
  }
  float whv[3] = {wh.x, wh.y, wh.z};
  *weight *= cvro_ggx_g1(alpha, wo, whv);
  return 1;
}


/* ------------------------------------------------------------------------- */
/* Optional decision / event log of ONE path (cvro_trace_path_logged): used by   */
/* the GPU parity tests to show that a path whose radiance differs from the      */
/* oracle's shares its event prefix with it up to one near-tie decision.         */
/* The log only OBSERVES: no arithmetic of the path loop changes.                */
/* ------------------------------------------------------------------------- */
static __thread cvro_trace* g_trace = 0;
static __thread const cvro_rng* g_trace_rng = 0;

static void log_decision(uint32_t kind, float a, float b) {
  cvro_trace* t = g_trace;
  if (!t) return;
  if (t->n_dec < t->cap_dec) {
    t->dec_kind[t->n_dec] = kind;
    t->dec_d[t->n_dec] = g_trace_rng->d;
    t->dec_a[t->n_dec] = a;
    t->dec_b[t->n_dec] = b;
  }
  t->n_dec++;
}
static void log_event(uint32_t code, uint32_t d) {
  cvro_trace* t = g_trace;
  if (!t) return;
  if (t->n_events < t->cap_events) {
    t->ev_code[t->n_events] = code;
    t->ev_d[t->n_events] = d;
  }
  t->n_events++;
}

/* ------------------------------------------------------------------------- */
/* Woodcock tracking (Utilities.cuh:129-155, Medium.h:135-143).  worldToAABB is  */
/* p - start/range (Q1).  One lookup past max_t is performed (Q10); the accept   */
/* draw is skipped when t > max_t.                                               */
/* ------------------------------------------------------------------------- */
static float woodcock(const cvro_scene* sc, v3 o, v3 d, float max_t, cvro_rng* rng,
                      cvro_counters* ctr) {
  float inv_max_sigmat = 1.0f / (sc->scale * sc->max_density);
  v3 bmin = V(sc->box_min[0], sc->box_min[1], sc->box_min[2]);
  v3 extent = vsub(V(sc->box_max[0], sc->box_max[1], sc->box_max[2]), bmin);
  v3 q = vdiv(bmin, extent);
  float event_density = 0.0f;
  float t = 0.0f;
  for (;;) {
    t += -logf(fmaxf(cvro_rng_float(rng), CVRO_EPS)) * inv_max_sigmat;
    v3 p = vsub(vadd(o, vscale(d, t)), q);
    float pp[3] = {p.x, p.y, p.z};
    event_density = sc->scale * cvro_density_lookup(sc, pp);
    ctr->density_lookups++;
    /* while (t <= max_t && event_density * inv_max_sigmat < u'): u' is drawn only when t <= max_t */
    log_decision(CVRO_DEC_EXIT, t, max_t);
    if (!(t <= max_t)) break;
    float u_accept = cvro_rng_float(rng);
    log_decision(CVRO_DEC_ACCEPT, event_density * inv_max_sigmat, u_accept);
    if (!(event_density * inv_max_sigmat < u_accept)) break;
  }
  return t;
}

/* unit pin of the above against the reference's own woodcockTracking (host-compiled,
 * oracle/ref_cpu_harness.cpp): Rng(seed), one call, returns t */
float cvro_woodcock(const cvro_scene* sc, const float o[3], const float d[3], float max_t,
                    int32_t seed, int* scattered) {
  cvro_rng r;
  cvro_counters c;
  memset(&c, 0, sizeof c);
  cvro_rng_init(&r, seed);
  float t = woodcock(sc, V(o[0], o[1], o[2]), V(d[0], d[1], d[2]), max_t, &r, &c);
  *scattered = t < max_t;
  return t;
}

/* ------------------------------------------------------------------------- */
/* the path loop (NaiveVolPTsk_kernel.cuh:22-86 /                                */
/* RegenerationVolPTsk_kernel.cuh:169-229)                                       */
/* ------------------------------------------------------------------------- */
int cvro_trace_path(const cvro_scene* sc, const cvro_camera* cam, cvro_rng* rng,
                    uint32_t image_id, int variant, uint32_t max_bounces,
                    float radiance[3], cvro_counters* ctr) {
  float of[3], df[3];
  float u0 = cvro_rng_float(rng);
  float u1 = cvro_rng_float(rng);
  cvro_camera_ray(cam, image_id, u0, u1, of, df);
  v3 o = V(of[0], of[1], of[2]), d = V(df[0], df[1], df[2]);
  float thr[3] = {1.f, 1.f, 1.f};
  float dist = 0.f, normal[3] = {0, 0, 0};
  int inside = 0;
  ctr->paths++;
  for (uint32_t bounce = 0;; ++bounce) {
    if (max_bounces && bounce >= max_bounces) return 0;
    ctr->bounces++;
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    if (!cvro_aabb_intersect(sc->box_min, sc->box_max, oo, dd, &dist, normal, &inside)) {
      /* Le == 1 (Medium.h:174-177) */
      radiance[0] = thr[0] * 1.f, radiance[1] = thr[1] * 1.f, radiance[2] = thr[2] * 1.f;
      ctr->escaped++;
      log_event(CVRO_EV_ESCAPE, rng->d);
      return 1;
    }
    int scattered = 0;
    float s = 0.f;
    if (inside) {
      s = woodcock(sc, o, d, dist, rng, ctr);
      scattered = s < dist;
    }
    const uint32_t ev_d = rng->d; /* draws consumed when the event starts */
    uint32_t ev_code = scattered ? CVRO_EV_SCATTER : CVRO_EV_BOUNDARY;
    if (!scattered) {
      float fx[3], fy[3], fz[3];
      cvro_frame_from_z(normal, fx, fy, fz);
      v3 X = V(fx[0], fx[1], fx[2]), Y = V(fy[0], fy[1], fy[2]), Z = V(fz[0], fz[1], fz[2]);
      v3 nd = vnormalize(V(-d.x, -d.y, -d.z));
      float wi[3] = {vdot(nd, X), vdot(nd, Y), vdot(nd, Z)};
      o = vadd(o, vscale(d, dist));
      float weight = 1;
      float wo[3] = {d.x, d.y, d.z}; /* *wo aliases path.ray.d in the reference */
      int ok;
      if (wi[2] == 0.f) {
        ok = 0;
      } else {
        float u[3];
        u[0] = cvro_rng_float(rng);
        u[1] = cvro_rng_float(rng);
        /* the third draw happens after the Fresnel term is known, but always happens */
        u[2] = cvro_rng_float(rng);
        int used;
        ok = cvro_ggx_sample(sc->ggx_alpha, sc->ggx_eta, wi, u, wo, &weight, &used);
      }
      /* on the failing reflect/refract checks the LOCAL wo has already been stored
       * into the ray direction (Bsdf.h:25-29 passes &path.ray.d) */
      d = V(wo[0], wo[1], wo[2]);
      ev_code |= (ok ? CVRO_EVF_OK : 0u) | (wo[2] < 0.f ? CVRO_EVF_WO_NEG : 0u) | (wi[2] < 0.f ? CVRO_EVF_WI_NEG : 0u);
      if (ok) {
        thr[0] *= weight, thr[1] *= weight, thr[2] *= weight;
        d = vadd(vadd(vscale(X, d.x), vscale(Y, d.y)), vscale(Z, d.z));
        o = vadd(o, vscale(d, CVRO_EPS));
      }
    } else {
      if (variant == CVRO_NAIVE)
        o = vsub(vadd(o, vscale(d, s)), vscale(d, CVRO_EPS));
      else
        o = vadd(o, vscale(d, s));
      v3 bmin = V(sc->box_min[0], sc->box_min[1], sc->box_min[2]);
      v3 bmax = V(sc->box_max[0], sc->box_max[1], sc->box_max[2]);
      v3 c = vdiv(vsub(o, bmin), vsub(bmax, bmin));
      float cc[3] = {c.x, c.y, c.z}, rgb[3];
      cvro_albedo_lookup(sc, cc, rgb);
      ctr->albedo_lookups++;
      thr[0] *= rgb[0], thr[1] *= rgb[1], thr[2] *= rgb[2];
      float e1 = cvro_rng_float(rng);
      float e2 = cvro_rng_float(rng);
      float dv[3] = {d.x, d.y, d.z}, nv[3];
      cvro_hg_sample(dv, sc->hg_g, e1, e2, nv);
      d = V(nv[0], nv[1], nv[2]);
    }
    /* Russian roulette (NaiveVolPTsk_kernel.cuh:75-84) */
    float p_survive = fminf(1.f, fmaxf(fmaxf(thr[0], thr[1]), thr[2]));
    if (fmaxf(fmaxf(thr[0], thr[1]), thr[2]) == 0.f) ev_code |= CVRO_EVF_ZERO;
    float u_rr = cvro_rng_float(rng);
    log_decision(CVRO_DEC_ROULETTE, u_rr, p_survive);
    if (u_rr > p_survive) {
      log_event(ev_code | CVRO_EVF_KILLED, ev_d);
      return 0;
    }
    log_event(ev_code, ev_d);
    thr[0] = thr[0] * 1.f / p_survive;
    thr[1] = thr[1] * 1.f / p_survive;
    thr[2] = thr[2] * 1.f / p_survive;
  }
}


int cvro_trace_path_logged(const cvro_scene* sc, const cvro_camera* cam, int32_t rng_seed,
                           uint32_t image_id, int variant, uint32_t max_bounces,
                           float radiance[3], cvro_trace* tr) {
  cvro_rng rng;
  cvro_counters ctr;
  memset(&ctr, 0, sizeof ctr);
  cvro_rng_init(&rng, rng_seed);
  tr->n_events = tr->n_dec = 0;
  tr->d0 = rng.d;
  g_trace = tr, g_trace_rng = &rng;
  radiance[0] = radiance[1] = radiance[2] = 0.f;
  int esc = cvro_trace_path(sc, cam, &rng, image_id, variant, max_bounces, radiance, &ctr);
  g_trace = 0, g_trace_rng = 0;
  return esc;
}

/* ------------------------------------------------------------------------- */
/* launch-level drivers                                                          */
/* ------------------------------------------------------------------------- */
typedef struct {
  const cvro_scene* sc;
  const cvro_camera* cam;
  uint64_t begin, end; /* pixel range [begin,end) for image jobs, path range for path jobs */
  uint32_t iterations;
  uint32_t seed;
  int mode; /* 0 naive image, 1 naive per-path, 2 regen path-rng, 3 regen thread-rng, 4 seeded per-path */
  int variant;
  uint64_t first;
  uint32_t n_persistent;
  uint64_t npix;
  float* out;
  cvro_counters ctr;
} job_t;

static void add_ctr(cvro_counters* a, const cvro_counters* b) {
  a->paths += b->paths;
  a->bounces += b->bounces;
  a->density_lookups += b->density_lookups;
  a->albedo_lookups += b->albedo_lookups;
  a->escaped += b->escaped;
}

static void* job_main(void* arg) {
  job_t* j = (job_t*)arg;
  memset(&j->ctr, 0, sizeof j->ctr);
  float rad[3];
  if (j->mode == 0 || j->mode == 2) {
    /* each host thread owns a pixel range: sums are in ascending sample order */
    int variant = j->mode == 0 ? CVRO_NAIVE : CVRO_REGEN;
    for (uint64_t p = j->begin; p < j->end; ++p) {
      for (uint32_t s = 0; s < j->iterations; ++s) {
        uint64_t path = (uint64_t)s * j->npix + p;
        cvro_rng rng;
        cvro_rng_init(&rng, (int32_t)(uint32_t)(j->seed + (uint32_t)path));
        if (cvro_trace_path(j->sc, j->cam, &rng, (uint32_t)p, variant, 0, rad, &j->ctr)) {
          float* px = j->out + 4 * p;
          px[0] += rad[0], px[1] += rad[1], px[2] += rad[2], px[3] = 1.f;
        }
      }
    }
  } else if (j->mode == 4) {
    /* per-path radiances with Rng(seed + path id): the path set of regenerationSK (rng=xorwow-path),
     * streamingMK (StreamingVolPTmk_kernel.cuh:55) and of the streaming kernels' per-path form */
    for (uint64_t path = j->begin; path < j->end; ++path) {
      cvro_rng rng;
      cvro_rng_init(&rng, (int32_t)(uint32_t)(j->seed + (uint32_t)path));
      float* px = j->out + 4 * (path - j->first);
      px[0] = px[1] = px[2] = px[3] = 0.f;
      if (cvro_trace_path(j->sc, j->cam, &rng, (uint32_t)(path % j->npix), j->variant, 0, rad, &j->ctr)) {
        px[0] = rad[0], px[1] = rad[1], px[2] = rad[2], px[3] = 1.f;
      }
    }
  } else if (j->mode == 1) {
    for (uint64_t path = j->begin; path < j->end; ++path) {
      cvro_rng rng;
      cvro_rng_init(&rng, (int32_t)(uint32_t)path);
      float* px = j->out + 4 * (path - j->seed /* first */);
      px[0] = px[1] = px[2] = px[3] = 0.f;
      if (cvro_trace_path(j->sc, j->cam, &rng, (uint32_t)(path % j->npix), CVRO_NAIVE, 0,
                          rad, &j->ctr)) {
        px[0] = rad[0], px[1] = rad[1], px[2] = rad[2], px[3] = 1.f;
      }
    }
  }
  return 0;
}

static void run_jobs(job_t* proto, uint64_t total, int n_host_threads,
                     cvro_counters* ctr, uint64_t base) {
  if (n_host_threads < 1) n_host_threads = 1;
  if ((uint64_t)n_host_threads > total && total > 0) n_host_threads = (int)total;
  job_t* jobs = (job_t*)calloc((size_t)n_host_threads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)n_host_threads, sizeof(pthread_t));
  for (int i = 0; i < n_host_threads; ++i) {
    jobs[i] = *proto;
    jobs[i].begin = base + total * (uint64_t)i / (uint64_t)n_host_threads;
    jobs[i].end = base + total * (uint64_t)(i + 1) / (uint64_t)n_host_threads;
    pthread_create(&th[i], 0, job_main, &jobs[i]);
  }
  cvro_counters sum;
  memset(&sum, 0, sizeof sum);
  for (int i = 0; i < n_host_threads; ++i) {
    pthread_join(th[i], 0);
    add_ctr(&sum, &jobs[i].ctr);
  }
  if (ctr) add_ctr(ctr, &sum);
  free(jobs);
  free(th);
}

static uint64_t tile_pixels(const cvro_camera* cam) {
  return (uint64_t)f2u_trunc(cam->resolution[0] * cam->resolution[1]);
}

void cvro_render_naive(const cvro_scene* sc, const cvro_camera* cam,
                       uint32_t iterations, float* out, int n_host_threads,
                       cvro_counters* ctr) {
  job_t j;
  memset(&j, 0, sizeof j);
  j.sc = sc, j.cam = cam, j.iterations = iterations, j.seed = 0, j.mode = 0;
  j.npix = tile_pixels(cam), j.out = out;
  run_jobs(&j, j.npix, n_host_threads, ctr, 0);
}

void cvro_trace_paths_naive(const cvro_scene* sc, const cvro_camera* cam,
                            uint64_t first, uint64_t count, float* out_per_path,
                            int n_host_threads, cvro_counters* ctr) {
  job_t j;
  memset(&j, 0, sizeof j);
  j.sc = sc, j.cam = cam, j.mode = 1, j.seed = (uint32_t)first;
  j.npix = tile_pixels(cam), j.out = out_per_path;
  run_jobs(&j, count, n_host_threads, ctr, first);
}

void cvro_trace_paths_seeded(const cvro_scene* sc, const cvro_camera* cam, uint64_t first,
                             uint64_t count, uint32_t seed, int variant, float* out_per_path,
                             int n_host_threads, cvro_counters* ctr) {
  job_t j;
  memset(&j, 0, sizeof j);
  j.sc = sc, j.cam = cam, j.mode = 4, j.seed = seed, j.variant = variant, j.first = first;
  j.npix = tile_pixels(cam), j.out = out_per_path;
  run_jobs(&j, count, n_host_threads, ctr, first);
}

/* thread-rng realisation of the reference queue: virtual thread v claims paths
 * v, v+T, v+2T, ... with ONE Rng(seed+v); run serially per virtual thread, the
 * virtual threads are split over host threads; per-pixel sums are then not in a
 * fixed order across host threads, so contributions are staged per path. */
typedef struct {
  const cvro_scene* sc;
  const cvro_camera* cam;
  uint32_t seed, T, v_begin, v_end;
  uint64_t n_paths, npix;
  float* staged; /* n_paths * 4 */
  cvro_counters ctr;
} vjob_t;

static void* vjob_main(void* arg) {
  vjob_t* j = (vjob_t*)arg;
  memset(&j->ctr, 0, sizeof j->ctr);
  for (uint32_t v = j->v_begin; v < j->v_end; ++v) {
    cvro_rng rng;
    cvro_rng_init(&rng, (int32_t)(j->seed + v));
    for (uint64_t path = v; path < j->n_paths; path += j->T) {
      float rad[3];
      float* px = j->staged + 4 * path;
      if (cvro_trace_path(j->sc, j->cam, &rng, (uint32_t)(path % j->npix), CVRO_REGEN, 0,
                          rad, &j->ctr)) {
        px[0] = rad[0], px[1] = rad[1], px[2] = rad[2], px[3] = 1.f;
        /* RegenerationVolPTsk_kernel.cuh:220-228: the roulette draw still happens
         * after an escape (Q8) */
        (void)cvro_rng_float(&rng);
      }
    }
  }
  return 0;
}

void cvro_render_regen(const cvro_scene* sc, const cvro_camera* cam,
                       uint32_t iterations, uint32_t seed, int rng_mode,
                       uint32_t n_persistent, float* out, int n_host_threads,
                       cvro_counters* ctr) {
  uint64_t npix = tile_pixels(cam);
  if (rng_mode == 1) {
    job_t j;
    memset(&j, 0, sizeof j);
    j.sc = sc, j.cam = cam, j.iterations = iterations, j.seed = seed, j.mode = 2;
    j.npix = npix, j.out = out;
    run_jobs(&j, npix, n_host_threads, ctr, 0);
    return;
  }
  uint64_t n_paths = npix * iterations;
  if (n_persistent < 1) n_persistent = 1;
  if (n_host_threads < 1) n_host_threads = 1;
  if ((uint32_t)n_host_threads > n_persistent) n_host_threads = (int)n_persistent;
  float* staged = (float*)calloc((size_t)n_paths * 4, sizeof(float));
  vjob_t* jobs = (vjob_t*)calloc((size_t)n_host_threads, sizeof(vjob_t));
  pthread_t* th = (pthread_t*)calloc((size_t)n_host_threads, sizeof(pthread_t));
  for (int i = 0; i < n_host_threads; ++i) {
    jobs[i].sc = sc, jobs[i].cam = cam, jobs[i].seed = seed, jobs[i].T = n_persistent;
    jobs[i].v_begin = (uint32_t)((uint64_t)n_persistent * i / n_host_threads);
    jobs[i].v_end = (uint32_t)((uint64_t)n_persistent * (i + 1) / n_host_threads);
    jobs[i].n_paths = n_paths, jobs[i].npix = npix, jobs[i].staged = staged;
    pthread_create(&th[i], 0, vjob_main, &jobs[i]);
  }
  for (int i = 0; i < n_host_threads; ++i) {
    pthread_join(th[i], 0);
    if (ctr) add_ctr(ctr, &jobs[i].ctr);
  }
  for (uint64_t path = 0; path < n_paths; ++path) {
    const float* s = staged + 4 * path;
    if (s[3] != 0.f) {
      float* px = out + 4 * (path % npix);
      px[0] += s[0], px[1] += s[1], px[2] += s[2], px[3] = 1.f;
    }
  }
  free(staged);
  free(jobs);
  free(th);
}
