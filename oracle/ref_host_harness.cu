/*
 * ref_host_harness.cu -- builds oracle/_ref/libcvr_ref_host.so.
 *
 * TEST INFRASTRUCTURE ONLY.  This TU contains NO reference code: it #includes the
 * reference's headers from where they lie under /root/reference (include path set
 * by oracle/Makefile) and exposes the reference's own __host__ __device__
 * functions, host-compiled through nvcc, behind a C ABI so that tests can pin
 * oracle/cvr_oracle.c against them and generate tests/golden/ vectors.
 *
 * Two accommodations, neither touching the reference tree:
 *  - an include dir with EMPTY glm/*.hpp files (written by the Makefile into
 *    oracle/_ref/shim) satisfies CVRMath.h:4-6, which needs no glm symbol;
 *  - RNG_H_ is pre-defined and a scripted `Rng` with the same interface is
 *    supplied, because the reference's Rng::getFloat() returns 0 on the host
 *    (Rng.h:24-30).
 */
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>

#include <helper_math.h>

#define RNG_H_
class Rng {
 public:
  const float* u_;
  int n_;
  __host__ __device__ Rng(const float* u = nullptr) : u_(u), n_(0) {}
  __host__ __device__ float getFloat() { return u_[n_++]; }
  __host__ __device__ float2 getFloat2() {
    float a = getFloat();
    float b = getFloat();
    return make_float2(a, b);
  }
  __host__ __device__ float3 getFloat3() {
    float a = getFloat();
    float b = getFloat();
    float c = getFloat();
    return make_float3(a, b, c);
  }
};

#include "Bsdf.h"
#include "CVRMath.h"
#include "GGX.h"
#include "Geometry.h"
#include "HG.h"
#include "Ray.h"
#include "Utilities.h"

extern "C" {

int ref_aabb_intersect(const float bmin[3], const float bmax[3], const float o[3],
                       const float d[3], float* dist, float normal[3], int* inside) {
  AABB box(make_float3(bmin[0], bmin[1], bmin[2]), make_float3(bmax[0], bmax[1], bmax[2]));
  SimpleIsect isect;
  isect.normal = make_float3(normal[0], normal[1], normal[2]);
  float3 ro = make_float3(o[0], o[1], o[2]);
  float3 rd = make_float3(d[0], d[1], d[2]);
  bool hit = box.intersect(ro, rd, isect);
  *dist = isect.dist;
  normal[0] = isect.normal.x, normal[1] = isect.normal.y, normal[2] = isect.normal.z;
  *inside = isect.inside_volume ? 1 : 0;
  return hit ? 1 : 0;
}

void ref_aabb_transform(const float bmin[3], const float bmax[3], float p[3]) {
  AABB box(make_float3(bmin[0], bmin[1], bmin[2]), make_float3(bmax[0], bmax[1], bmax[2]));
  float3 q = make_float3(p[0], p[1], p[2]);
  box.transform(q);
  p[0] = q.x, p[1] = q.y, p[2] = q.z;
}

void ref_frame_from_z(const float n[3], float x[3], float y[3], float z[3]) {
  Frame f;
  f.setFromZ(make_float3(n[0], n[1], n[2]));
  x[0] = f.x_.x, x[1] = f.x_.y, x[2] = f.x_.z;
  y[0] = f.y_.x, y[1] = f.y_.y, y[2] = f.y_.z;
  z[0] = f.z_.x, z[1] = f.z_.y, z[2] = f.z_.z;
}

void ref_frame_local_world(const float n[3], const float a[3], float local[3], float world[3]) {
  Frame f;
  f.setFromZ(make_float3(n[0], n[1], n[2]));
  float3 l = f.toLocal(make_float3(a[0], a[1], a[2]));
  float3 w = f.toWorld(make_float3(a[0], a[1], a[2]));
  local[0] = l.x, local[1] = l.y, local[2] = l.z;
  world[0] = w.x, world[1] = w.y, world[2] = w.z;
}

void ref_hg_sample(const float dir[3], float g, float e1, float e2, float out[3]) {
  float3 r = ImportanceSampleHG(make_float3(dir[0], dir[1], dir[2]), g, e1, e2);
  out[0] = r.x, out[1] = r.y, out[2] = r.z;
}

int ref_ggx_sample(const float alpha[2], float eta, const float wi[3], const float u[3],
                   float wo[3], float* weight, int* n_used) {
  Rng rng(u);
  float3 out = make_float3(wo[0], wo[1], wo[2]);
  bool ok = GGX_sample(make_float2(alpha[0], alpha[1]), eta,
                       make_float3(wi[0], wi[1], wi[2]), &rng, &out, weight);
  wo[0] = out.x, wo[1] = out.y, wo[2] = out.z;
  *n_used = rng.n_;
  return ok ? 1 : 0;
}

void ref_ggx_defaults(float alpha[2], float* eta) {
  GGX g;
  alpha[0] = g.roughness.x, alpha[1] = g.roughness.y;
  *eta = g.int_ior_over_ext_ior;
}

float ref_fresnel_dielectric(float eta, float ndotwi, float* ndotwt) {
  return fresnelDielectric(eta, ndotwi, ndotwt);
}

float ref_ggx_g1(const float alpha[2], const float v[3], const float m[3]) {
  return GGX_G1(make_float2(alpha[0], alpha[1]), make_float3(v[0], v[1], v[2]),
                make_float3(m[0], m[1], m[2]));
}

unsigned int ref_morton3d(float x, float y, float z) { return morton3D(x, y, z); }

float ref_scale(float x, float s) { return UtilityFunctors::Scale(s)(x); }

float ref_fmaxf3(float x, float y, float z) { return fmaxf3(make_float4(x, y, z, 0.f)); }

}  // extern "C"
