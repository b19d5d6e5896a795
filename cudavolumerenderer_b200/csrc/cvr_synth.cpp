// cvr_synth.cpp -- procedural stand-ins for the reference's scenes (host code).
//
// Every volume payload in the reference checkout is a Git-LFS pointer stub
// (SURVEY.md section 2 row 18), so the named scenes are synthesised with the SHAPES
// and medium parameters of the originals:
//   "bucky"  32^3 uint8 phantom pushed through the Raw loader's arithmetic
//            (RawSceneBuilder.h:35-83 density = byte/max; :95-139 transfer function)
//   "hetvol" 128x128x50 smoke-like fBm density, constant albedo (0.96,0.84,0.68)
//   "manix"  256x230x256 head phantom through the MHD->VDB converter's maths
//            (scripts/convert-mhd/mhd_to_vdb.py:47-63: normalise, smoothstep(0.2,0.6),
//            albedo = (rho, 0, 0))
//   "fbm"    n^3 value-noise fBm, rho = max(0, fbm-0.4)/0.6
// Deterministic for a given seed; multithreaded over z slices.
#include "../../include/cvr_abi.h"
#include "cvr_noise.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

using cvrnoise::fbm5;

template <class F>
void parallel_z(int nz, F f) {
  unsigned nt = std::max(1u, std::min((unsigned)nz, std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([=]() {
      for (int z = (int)t; z < nz; z += (int)nt) f(z);
    });
  for (auto& x : th) x.join();
}

inline float smoothstep(float e0, float e1, float x) {
  float t = std::min(std::max((x - e0) / (e1 - e0), 0.0f), 1.0f);
  return t * t * (3.0f - 2.0f * t);
}

// RawSceneBuilder.h:95-139 behaviour: two linear ramps sampled at 20 + 80 points
std::vector<float> transfer_function() {
  std::vector<float> tf;
  const float len = 100.f;
  float sr = 0.02f, sg = 0.2f, sb = 0.02f, er = 1.f, eg = 0.02f, eb = 0.02f;
  for (int i = 0; i < len * 1.f / 5.f; i++) {
    tf.push_back(sr + (i * (er - sr) / len));
    tf.push_back(sg + (i * (eg - sg) / len));
    tf.push_back(sb + (i * (eb - sb) / len));
  }
  sr = er, sg = eg, sb = eb;
  er = 0.0f, eg = 0.02f, eb = 1.0f;
  for (int i = 0; i < len * 4.f / 5.f; i++) {
    tf.push_back(sr + (i * (er - sr) / len));
    tf.push_back(sg + (i * (eg - sg) / len));
    tf.push_back(sb + (i * (eb - sb) / len));
  }
  return tf;
}

}  // namespace

extern "C" int cvr_synth_volume(const char* kind_c, int32_t nx, int32_t ny, int32_t nz, uint32_t seed,
                                float* density, float* albedo, float* max_density) {
  if (!kind_c || !density || nx < 2 || ny < 2 || nz < 2) return 1;
  const std::string kind(kind_c);
  const size_t n = (size_t)nx * ny * nz;
  if (kind == "bucky") {
    // uint8 phantom, then the Raw loader's normalisation
    std::vector<unsigned char> raw(n);
    parallel_z(nz, [&](int z) {
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
          float px = ((x + .5f) / nx - .5f) * 2.f, py = ((y + .5f) / ny - .5f) * 2.f,
                pz = ((z + .5f) / nz - .5f) * 2.f;
          float r = std::sqrt(px * px + py * py + pz * pz);
          float shell = std::exp(-((r - 0.62f) / 0.12f) * ((r - 0.62f) / 0.12f));
          float pat = 0.5f + 0.5f * std::cos(6.f * std::atan2(py, px)) *
                                 std::cos(6.f * std::acos(pz / std::max(r, 1e-6f)));
          float v = std::floor(255.f * shell * pat + 0.5f);
          raw[x + (size_t)nx * (y + (size_t)ny * z)] = (unsigned char)std::min(std::max(v, 0.f), 255.f);
        }
    });
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) {
      density[i] = raw[i];
      mx = std::fmax(density[i], mx);
    }
    if (mx <= 0.f) mx = 1.f;
    for (size_t i = 0; i < n; ++i) density[i] /= mx;
    if (albedo) {
      std::vector<float> tf = transfer_function();
      size_t entries = tf.size() / 3;
      for (size_t i = 0; i < n; ++i) {
        float v = density[i] * (entries - 1);
        size_t k = (size_t)std::ceil(v);
        albedo[4 * i + 0] = tf[3 * k], albedo[4 * i + 1] = tf[3 * k + 1], albedo[4 * i + 2] = tf[3 * k + 2];
        albedo[4 * i + 3] = 1.f;
      }
    }
    if (max_density) *max_density = 1.f;  // RawSceneBuilder.h:68
    return 0;
  }
  if (kind == "hetvol") {
    const uint32_t sd = seed ? seed : 0x5e0du;
    parallel_z(nz, [&](int z) {
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
          float qx = (x + .5f) / nx - .5f, qy = (y + .5f) / ny - .5f, qz = (z + .5f) / nz - .5f;
          float f = fbm5(8.f * (qx + .5f), 8.f * (qy + .5f), 8.f * (qz + .5f), sd);
          float rho = f - 0.35f * (1.f + std::sqrt(qx * qx + qy * qy));
          rho = std::min(std::max(rho, 0.f), 1.f);
          density[x + (size_t)nx * (y + (size_t)ny * z)] = rho;
        }
    });
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) mx = std::max(std::min(1.0f, density[i]), mx);  // XmlSceneBuilder.h:187
    if (albedo)
      for (size_t i = 0; i < n; ++i) {
        albedo[4 * i + 0] = 0.96f, albedo[4 * i + 1] = 0.84f, albedo[4 * i + 2] = 0.68f, albedo[4 * i + 3] = 1.f;
      }
    if (max_density) *max_density = mx > 0.f ? mx : 1.f;
    return 0;
  }
  if (kind == "manix") {
    const uint32_t sd = seed ? seed : 0x3a91u;
    parallel_z(nz, [&](int z) {
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
          float ux = ((x + .5f) / nx - .5f) * 2.f, uy = ((y + .5f) / ny - .5f) * 2.f,
                uz = ((z + .5f) / nz - .5f) * 2.f;
          float e_skin = (ux / 0.85f) * (ux / 0.85f) + (uy / 0.95f) * (uy / 0.95f) + (uz / 0.9f) * (uz / 0.9f);
          float e_bone = (ux / 0.78f) * (ux / 0.78f) + (uy / 0.88f) * (uy / 0.88f) + (uz / 0.83f) * (uz / 0.83f);
          float e_brain = (ux / 0.70f) * (ux / 0.70f) + (uy / 0.80f) * (uy / 0.80f) + (uz / 0.75f) * (uz / 0.75f);
          float v = 0.f;
          if (e_skin < 1.f) v = 0.27f;
          if (e_bone < 1.f) v = 1.0f;
          if (e_brain < 1.f) {
            v = 0.33f + 0.12f * fbm5(6.f * (ux + 1.f), 6.f * (uy + 1.f), 6.f * (uz + 1.f), sd);
            // a few vessels: tubes around sinusoidal centre lines
            for (int k = 0; k < 3; ++k) {
              float cy = 0.35f * std::sin(3.f * ux + 2.1f * k), cz = 0.3f * std::cos(2.f * ux + 1.3f * k);
              float d2 = (uy - cy) * (uy - cy) + (uz - cz) * (uz - cz);
              if (d2 < 0.0016f) v = 0.7f;
            }
          }
          density[x + (size_t)nx * (y + (size_t)ny * z)] = v;
        }
    });
    float lo = density[0], hi = density[0];
    for (size_t i = 0; i < n; ++i) lo = std::min(lo, density[i]), hi = std::max(hi, density[i]);
    float mx = 0.f;
    for (size_t i = 0; i < n; ++i) {
      float nrm = (density[i] - lo) / (hi - lo);   // mhd_to_vdb.py:47-50
      float rho = smoothstep(0.2f, 0.6f, nrm);     // :51-53
      density[i] = rho;
      mx = std::max(mx, rho);
      if (albedo) {
        albedo[4 * i + 0] = rho, albedo[4 * i + 1] = 0.f, albedo[4 * i + 2] = 0.f, albedo[4 * i + 3] = 1.f;  // :61-63
      }
    }
    if (max_density) *max_density = mx;  // VDBSceneBuilder.h:54-55
    return 0;
  }
  if (kind == "fbm" || kind == "sparsefbm") {
    const uint32_t sd = seed ? seed : 0x5eedu;
    const bool sparse = kind == "sparsefbm";
    std::vector<float> zmax((size_t)nz, 0.f);
    parallel_z(nz, [&](int z) {
      float m = 0.f;
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
          float rho = sparse ? cvrnoise::sparsefbm_density(x, y, z, sd) : cvrnoise::fbm_density(x, y, z, sd);
          density[x + (size_t)nx * (y + (size_t)ny * z)] = rho;
          m = std::max(m, rho);
        }
      zmax[z] = m;
    });
    float mx = *std::max_element(zmax.begin(), zmax.end());
    if (albedo)
      for (size_t i = 0; i < n; ++i) {
        albedo[4 * i + 0] = albedo[4 * i + 1] = albedo[4 * i + 2] = 0.99f;
        albedo[4 * i + 3] = 1.f;
      }
    if (max_density) *max_density = mx > 0.f ? mx : 1.f;
    return 0;
  }
  return 1;
}
