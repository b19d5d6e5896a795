// cvr_kernels.cuh -- the sm_100a path kernels.
//
// Persistent-thread kernels replace the reference's naiveSK / regenerationSK(thread) /
// streamingSK / streamingMK / sortingSK kernels (NaiveVolPTsk_kernel.cuh:17-87,
// RegenerationVolPTsk_kernel.cuh:146-232, StreamingVolPTsk_kernel.cuh:219-290, ...): the kernel
// NAME selects the reference SEMANTICS (scatter pull-back, seed advance, roulette-after-escape),
// the scheduling is always this library's.  Four schedulers share the estimator blocks below
// (do_isect, do_track_step / track_pair_fast, do_scatter, do_boundary, do_roulette):
//
//  * k_volpt_warp    (default) warp-private wavefront: every warp owns 64/96 path slots in shared
//                    memory and runs batches of <= 32 paths in the SAME state; no atomics;
//  * k_volpt_queued  per-state ring-buffer queues per CTA (the first-round product kernel);
//  * k_volpt_sorted  block-wide counting sort per round;
//  * k_volpt         a lane keeps its path in registers ("while-while" with a yield threshold).
//
// Common to all: grid = SMs x resident CTAs, warps loop until the 64-bit path counter is
// exhausted; idle lanes claim path ids with ONE warp-aggregated atomic (ballot + popc +
// shuffle), consecutive ids -> consecutive pixels; a density lookup reads one 32-byte cell
// (8 trilinear corners) with a single 256-bit load, an albedo lookup one 128-byte line; a
// path's own operation order and RNG draw order never depend on the scheduler.
#pragma once
#include "cvr_device.cuh"

namespace cvr {

enum : int { RNG_XORWOW_PATH = 0, RNG_XORWOW_THREAD = 1, RNG_PHILOX = 2 };
// LAYOUT_BRICK: the cell8 cells of LAYOUT_CELL8 stored SPARSELY -- bricks of 8^3 cells (16 KB),
// only bricks that can hold a non-zero value are kept, a dense table over the brick grid maps
// a brick coordinate to its slot (slot 0 = one shared all-zero brick).  A lookup costs one
// extra dependent 4-byte load (the table is small enough to stay in L2).
enum : int { LAYOUT_LINEAR = 0, LAYOUT_CELL8 = 1, LAYOUT_BRICK = 2 };

struct DeviceCounters {
  unsigned long long paths, bounces, density_lookups, albedo_lookups, escaped, speculative, skipped;
};

struct TrackInv {
  float inv_max_sigmat;  // 1 / (scale * max_density)
  float qx, qy, qz;      // box_min / (box_max - box_min)
  float rx, ry, rz;      // (float)(uint)(res - 1)
  uint32_t nx, ny, nz;   // density dims
  uint32_t sy, sz;       // cell strides (cell8) in cells
  // fused ("exact=0") forms
  float nqrx, nqry, nqrz;  // -q * r : grid coordinate = fma(p, r, nqr)
  float sig_ratio;         // scale * inv_max_sigmat : accept test = density * sig_ratio < u
  float aix, aiy, aiz;     // 1 / (box_max - box_min) for the albedo coordinate
  float neg_ln2_inv_sigmat;  // -ln(2) * inv_max_sigmat : free-flight step = lg2(u) * this
  // local-majorant tracking ("tracking=local"): max density per brick of CVR_BRICK^3 cells
  const float* __restrict__ majorant;
  uint32_t mx, my, mz;     // brick-grid dims = ceil((n + 1) / CVR_BRICK)
  // second majorant level: max over 8^3 bricks = 64^3 cells (empty space is crossed in 64-cell strides)
  const float* __restrict__ majorant2;
  uint32_t m2x, m2y, m2z;
};

#ifndef CVR_BRICK_LOG2
#define CVR_BRICK_LOG2 3
#endif
#define CVR_BRICK (1 << CVR_BRICK_LOG2)

struct KernelParams {
  CameraParams cam;
  MediumParams med;
  TrackInv inv;
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
  uint2* path_log;      // debug: per-path event log, log_cap entries of (code, draw counter) per path (or nullptr)
  uint32_t log_cap;
  unsigned long long* head;  // path queue head
  DeviceCounters* ctr;
  uint32_t max_bounces;
  int loop_threshold;
  int pullback;         // naive/streaming: o -= d*EPSILON at scatter (Q8)
  int rr;               // Russian roulette enabled (Defines.h:44)
  int rr_after_escape;  // thread-rng regeneration/streaming: the roulette draw is consumed after an escape (Q8)
  int track_steps;      // sorted scheduler: Woodcock steps per round
  int track_min_lanes;  // sorted scheduler: leave the step loop when fewer lanes are still tracking
  int fix_nan;          // drop non-finite path contributions (reference quirk opt-out, default 0)
  int pair;             // fast tracking loop: two Woodcock steps per iteration, the second speculative
  int policy;           // warp scheduler: 0 = fullest state wins, 1 = events first unless a full tracking batch waits
  int exit_others;      // warp scheduler: apply track_min_lanes also when this many other slots of the warp wait (0 = tracking slots only)
  // fetch-skip table (warp scheduler, fused arithmetic, global majorant; see SkipTab below)
  const uint8_t* skip_tab;  // one byte per brick of (1 << skip_shift)^3 lookup cells, x fastest
  uint32_t skip_n;          // bytes in the table (0 = feature off)
  uint32_t skip_shift, skip_bx, skip_bxy;
  uint32_t regen_block;  // 1: path ids walk 8 x 4 pixel blocks of the tile (regen_order=block)  // brick edge = 1 << shift cells; row / slice strides in bricks
};

// S_BOUNDARY_P = boundary event whose FIRST uniform is already drawn and parked in
// PathRegs::t (the fast tracking loop draws both uniforms of a Woodcock step up front; the
// reference does not consume the second one when the step left the medium).
enum : int { S_IDLE = 0, S_ISECT = 1, S_TRACK = 2, S_SCATTER = 3, S_BOUNDARY = 4, S_DONE = 5, S_BOUNDARY_P = 6 };

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
CVR_DEV V3 albedo_lookup(const MediumParams& m, V3 p) {
  if (m.albedo_const) return v3(m.albedo_r, m.albedo_g, m.albedo_b);
  if (LAYOUT != LAYOUT_LINEAR) return albedo_cell8(m, p);
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

// Per-path registers.  One lane owns one path at a time.
template <class Rng>
struct PathRegs {
  V3 o, d;
  float thr_x, thr_y, thr_z;
  float t, dist;
  Rng rng;
  uint32_t out_idx;
  uint32_t path_lo;  // index of the path inside this launch (per-path debug output)
  uint32_t bounces;
  int ncode;
  int state;
};

struct LaneCounters {
  uint32_t paths = 0, bounces = 0, dens = 0, alb = 0, esc = 0, spec = 0, skip = 0;
  // pair loop: iterations and iterations whose second step really happened; folded into
  // dens (= pairs + cont) and spec (= pairs - cont) when the counters are flushed
  uint32_t pairs = 0, cont = 0;
};

// Loop invariants of the Woodcock step (the reference recomputes them every
// iteration: Utilities.cuh:143, 129-132; Volume.h:40-45).  They are computed ONCE on
// the host with the same fp32 operations (cvr_abi.cu: fill_track_inv) and live in the
// kernel-parameter constant bank, so they cost neither registers nor instructions.
// density at normalised coordinate p (A7) with the invariants hoisted; same values,
// same operation order as density_cell8 / density_linear in cvr_device.cuh
// float offset of cell (kx,ky,kz) inside m.dcells
template <int LAYOUT>
CVR_DEV size_t cell_offset(const MediumParams& m, const TrackInv& I, uint32_t kx, uint32_t ky, uint32_t kz) {
  if (LAYOUT == LAYOUT_BRICK) {
    const uint32_t b = (kx >> 3) + m.bmx * ((ky >> 3) + m.bmy * (kz >> 3));
    const uint32_t slot = __ldg(m.btable + b);
    return ((size_t)slot * 512u + ((kx & 7u) | ((ky & 7u) << 3) | ((kz & 7u) << 6))) * 8u;
  }
  // cells < 2^32 is guaranteed by cvr_set_scene
  return 8 * (size_t)(kx + I.sy * ky + I.sz * kz);
}

template <int LAYOUT>
CVR_DEV float density_at(const MediumParams& m, const TrackInv& I, V3 p) {
  if (LAYOUT == LAYOUT_LINEAR) return density_linear(m, p);
  float cx = p.x * I.rx, cy = p.y * I.ry, cz = p.z * I.rz;
  int x1 = floorf(cx), y1 = floorf(cy), z1 = floorf(cz);
  float fx = cx - x1, fy = cy - y1, fz = cz - z1;
  uint32_t kx = min((uint32_t)(x1 + 1), I.nx), ky = min((uint32_t)(y1 + 1), I.ny),
           kz = min((uint32_t)(z1 + 1), I.nz);
  float v[8];
  ldg256_d(m.dcells + cell_offset<LAYOUT>(m, I, kx, ky, kz), v);
  return trilerp<true>(v[0], v[2], v[4], v[6], v[1], v[3], v[5], v[7], fx, fy, fz);
}

// ---- fused-arithmetic ("exact=0") building blocks -------------------------------------
// Same algorithm, same RNG draws in the same order; only the floating-point evaluation
// differs from the reference's: MUFU-based log / reciprocal / rsqrt / sincos instead of
// the IEEE-rounded library calls, fma-folded coordinate transforms, a+f*(b-a) lerps.
// Results agree with the exact mode to ~1e-6 relative per operation; parity of this
// mode is the statistical one.
CVR_DEV float lerp_fast(float a, float b, float f) { return fmaf(f, b - a, a); }
#ifndef CVR_TRILERP_X2
#define CVR_TRILERP_X2 1
#endif
// Blackwell packed fp32 arithmetic (PTX add/sub/fma .f32x2 -> SASS FADD2 / FFMA2): one
// instruction blends two corner pairs.
typedef unsigned long long f32x2_t;
CVR_DEV f32x2_t pack2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
CVR_DEV void unpack2(f32x2_t r, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
CVR_DEV f32x2_t lerp2(f32x2_t a, f32x2_t b, f32x2_t f) {  // a + f * (b - a), both halves
  f32x2_t d, r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(b), "l"(a));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f), "l"(d), "l"(a));
  return r;
}
// v in the cell8 corner order (cvr_device.cuh): pairs (v0,v1)=(x1,y1,z1|z2), (v2,v3)=(x2,y1),
// (v4,v5)=(x1,y2), (v6,v7)=(x2,y2): x blend, y blend on pairs, z blend scalar = 8 instructions
CVR_DEV float trilerp_fast(const float (&v)[8], float fx, float fy, float fz) {
#if CVR_TRILERP_X2
  const f32x2_t FX = pack2(fx, fx), FY = pack2(fy, fy);
  f32x2_t R = lerp2(pack2(v[0], v[1]), pack2(v[2], v[3]), FX);
  f32x2_t S = lerp2(pack2(v[4], v[5]), pack2(v[6], v[7]), FX);
  float t0, t1;
  unpack2(lerp2(R, S, FY), t0, t1);
  return lerp_fast(t0, t1, fz);
#else
  float r0 = lerp_fast(v[0], v[2], fx), r1 = lerp_fast(v[1], v[3], fx);
  float s0 = lerp_fast(v[4], v[6], fx), s1 = lerp_fast(v[5], v[7], fx);
  return lerp_fast(lerp_fast(r0, s0, fy), lerp_fast(r1, s1, fy), fz);
#endif
}
// floor() without the conversion pipe: x + 1.5*2^23 rounded toward -inf is floor(x) + 1.5*2^23
// exactly for |x| < 2^22, so the integer is in the low mantissa bits and the float floor is
// one exact subtraction away (FADD.RM + FADD instead of F2I.FLOOR + FRND.FLOOR).
#define CVR_FLOOR_MAGIC 12582912.0f
#define CVR_FLOOR_MAGIC_BITS 0x4B400000
// A lookup split into "address + issue the 256-bit load" and "blend", so that a caller can
// put several loads in flight before the first blend.
struct CellFetch {
  float v[8];
  float fx, fy, fz;
};
template <int LAYOUT>
CVR_DEV void cell_fetch(const MediumParams& m, const TrackInv& I, float cx, float cy, float cz, CellFetch& F) {
  const float bx = __fadd_rd(cx, CVR_FLOOR_MAGIC), by = __fadd_rd(cy, CVR_FLOOR_MAGIC), bz = __fadd_rd(cz, CVR_FLOOR_MAGIC);
  // cell index k = x1 + 1, clamped like the reference's fetches (Q2: negative x1 wraps to the far edge)
  uint32_t kx = min((uint32_t)(__float_as_int(bx) - (CVR_FLOOR_MAGIC_BITS - 1)), I.nx),
           ky = min((uint32_t)(__float_as_int(by) - (CVR_FLOOR_MAGIC_BITS - 1)), I.ny),
           kz = min((uint32_t)(__float_as_int(bz) - (CVR_FLOOR_MAGIC_BITS - 1)), I.nz);
  ldg256_d(m.dcells + cell_offset<LAYOUT>(m, I, kx, ky, kz), F.v);
  F.fx = cx - (bx - CVR_FLOOR_MAGIC), F.fy = cy - (by - CVR_FLOOR_MAGIC), F.fz = cz - (bz - CVR_FLOOR_MAGIC);
}
// ---- fetch-skip table ----------------------------------------------------------------------
// The Woodcock loop is bound by the L1TEX wavefront rate, not by bytes or issue slots: every
// lane of a warp looks up a different 32-byte cell, so one LDG.256 is ~21 wavefronts, and the
// random-sector gather microbenchmark tops out at ~0.4 sectors / cycle / SM -- the rate this
// kernel already runs at.  The only way past it is to fetch less.  A step whose
// accept draw w exceeds (brick majorant) * sig_ratio is a null collision WHATEVER the cell
// holds: density <= majorant, the multiply is monotonic, so `density * sig_ratio < w` is
// certain.  The cell load of such a step is skipped.  RNG consumption, the accept decision and
// therefore every path are unchanged bit for bit (unlike tracking=local, which changes the
// sequence); only the number of gathers drops (hetvol -53 %, manix -90 %).
// The majorant lives in SHARED memory (a global table would cost the same wavefront as the
// cell): one byte per brick of (1 << shift)^3 cells, s = ceil(r * 256 * (1 + 1e-5)) - 1 clamped
// to [0, 255] with r = majorant * sig_ratio in [0, 1]; the test is on the TOP BYTE of the raw
// XORWOW word behind w (w > top / 256): top > s  =>  w > (s + 1) / 256 >= r.  The 1e-5 margin
// covers the few ulps a fused trilinear blend can exceed the largest corner by; s = 255 never
// skips.  Skipped lanes load cell 0 instead (one line shared by all of them).
// The geometry (shift, strides) stays in the kernel-parameter constant bank and the table sits at
// a compile-time offset of the dynamic shared memory, so the test costs no registers.
struct SkipTab {
  const uint8_t* tab;  // shared memory
};
template <bool SKIP>
CVR_DEV bool skip_test(const KernelParams& P, const SkipTab& S, uint32_t kx, uint32_t ky, uint32_t kz, uint32_t w_word) {
  if (!SKIP) return false;
  const uint32_t b = (kx >> P.skip_shift) + (ky >> P.skip_shift) * P.skip_bx + (kz >> P.skip_shift) * P.skip_bxy;
  return (w_word >> 24) > (uint32_t)S.tab[b];
}
// cell_fetch with the skip test between the address and the load
template <int LAYOUT, bool SKIP>
CVR_DEV bool cell_fetch_skip(const KernelParams& P, const TrackInv& I, const SkipTab& S, float cx, float cy, float cz,
                             uint32_t w_word, CellFetch& F) {
  const MediumParams& m = P.med;
  const float bx = __fadd_rd(cx, CVR_FLOOR_MAGIC), by = __fadd_rd(cy, CVR_FLOOR_MAGIC), bz = __fadd_rd(cz, CVR_FLOOR_MAGIC);
  uint32_t kx = min((uint32_t)(__float_as_int(bx) - (CVR_FLOOR_MAGIC_BITS - 1)), I.nx),
           ky = min((uint32_t)(__float_as_int(by) - (CVR_FLOOR_MAGIC_BITS - 1)), I.ny),
           kz = min((uint32_t)(__float_as_int(bz) - (CVR_FLOOR_MAGIC_BITS - 1)), I.nz);
  const bool skip = skip_test<SKIP>(P, S, kx, ky, kz, w_word);
  // a skipped lane loads cell 0 instead: any resident cell will do, the caller ignores the blend.
  // (Predicating the lane off the load was measured and rejected: it does save the lane's L1
  // data-pipe wavefront -- the pipe charges one per ACTIVE lane of an LDG.256 wherever it points --
  // but the destination registers then need a definition (4 CS2R per load), and the ~30 extra
  // instructions per pair cost more than the wavefronts: manix +5.7 % instead of +8.6 %,
  // fbm 512^3 +31 % instead of +37 % over skip=0.)
  if (SKIP && skip) kx = ky = kz = 0u;
  ldg256_d<SKIP>(m.dcells + cell_offset<LAYOUT>(m, I, kx, ky, kz), F.v);
  F.fx = cx - (bx - CVR_FLOOR_MAGIC), F.fy = cy - (by - CVR_FLOOR_MAGIC), F.fz = cz - (bz - CVR_FLOOR_MAGIC);
  return skip;
}
template <int LAYOUT>
CVR_DEV float density_at_grid(const MediumParams& m, const TrackInv& I, float cx, float cy, float cz) {
  CellFetch F;
  cell_fetch<LAYOUT>(m, I, cx, cy, cz, F);
  return trilerp_fast(F.v, F.fx, F.fy, F.fz);
}
template <int LAYOUT>
CVR_DEV float density_at_fast(const MediumParams& m, const TrackInv& I, V3 p) {
  return density_at_grid<LAYOUT>(m, I, fmaf(p.x, I.rx, I.nqrx), fmaf(p.y, I.ry, I.nqry), fmaf(p.z, I.rz, I.nqrz));
}
CVR_DEV float lg2_fast(float x) {  // x >= 1e-5 here: no denormal fix-up needed
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
CVR_DEV V3 albedo_cell8_fast(const MediumParams& m, V3 p) {
  float cx = p.x * (float)(uint32_t)(m.anx - 1), cy = p.y * (float)(uint32_t)(m.any - 1),
        cz = p.z * (float)(uint32_t)(m.anz - 1);
  float flx = floorf(cx), fly = floorf(cy), flz = floorf(cz);
  size_t kx = cell_index((int)flx, m.anx), ky = cell_index((int)fly, m.any), kz = cell_index((int)flz, m.anz);
  AlbedoCell a;
  ldg_albedo_cell(m.acells + CVR_ACELL_FLOATS * (kx + (size_t)(m.anx + 1) * (ky + (size_t)(m.any + 1) * kz)), a);
  float fx = cx - flx, fy = cy - fly, fz = cz - flz;
  V3 out;
#if CVR_TRILERP_X2
  // (r, g) pairs of corner i = x | y << 1 | z << 2: x, y, z blends on packed pairs; b as a density cell
  const f32x2_t FX = pack2(fx, fx), FY = pack2(fy, fy), FZ = pack2(fz, fz);
  f32x2_t rg[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    rg[i] = lerp2(pack2(a.rg[4 * i], a.rg[4 * i + 1]), pack2(a.rg[4 * i + 2], a.rg[4 * i + 3]), FX);
  unpack2(lerp2(lerp2(rg[0], rg[1], FY), lerp2(rg[2], rg[3], FY), FZ), out.x, out.y);
  out.z = trilerp_fast(a.b, fx, fy, fz);
#else
  float r[8], g[8];
  // trilerp_fast expects the density corner order k = z | x << 1 | y << 2
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int i = ((k >> 1) & 1) | (((k >> 2) & 1) << 1) | ((k & 1) << 2);
    r[k] = a.rg[2 * i], g[k] = a.rg[2 * i + 1];
  }
  out = v3(trilerp_fast(r, fx, fy, fz), trilerp_fast(g, fx, fy, fz), trilerp_fast(a.b, fx, fy, fz));
#endif
  return out;
}
// Same test, but the hit face comes out as the 0..5 normal code the slots store (0:+x 1:+y 2:+z
// 3:-x 4:-y 5:-z) and `inside` from the sign of that one direction component, instead of a
// float3 normal, a dot product and a second compare chain.  No face matched (NaN): code 5,
// inside false -- what code_from_normal / dot give for the zero normal.
CVR_DEV bool box_intersect_code(const V3& bmin, const V3& bmax, const V3& o, const V3& d, float& dist, int& code,
                                bool& inside) {
  V3 inv_r = v3(__fdividef(1.0f, d.x), __fdividef(1.0f, d.y), __fdividef(1.0f, d.z));
  V3 tbot = inv_r * (bmin - o);
  V3 ttop = inv_r * (bmax - o);
  V3 tmin = v3(fminf(ttop.x, tbot.x), fminf(ttop.y, tbot.y), fminf(ttop.z, tbot.z));
  V3 tmax = v3(fmaxf(ttop.x, tbot.x), fmaxf(ttop.y, tbot.y), fmaxf(ttop.z, tbot.z));
  float largest_tmin = fmaxf(fmaxf(tmin.x, tmin.y), fmaxf(tmin.x, tmin.z));
  float smallest_tmax = fminf(fminf(tmax.x, tmax.y), fminf(tmax.x, tmax.z));
  dist = (largest_tmin > CVR_EPS) ? largest_tmin : smallest_tmax;
  // Geometry.h:76-87: first match in the order top x, y, z, bottom x, y, z
  code = dist == ttop.x ? 0 : dist == ttop.y ? 1 : dist == ttop.z ? 2 : dist == tbot.x ? 3 : dist == tbot.y ? 4 : 5;
  const bool matched = code < 5 || dist == tbot.z;
  const float dn = code == 0 ? d.x : code == 1 ? d.y : code == 2 ? d.z : code == 3 ? -d.x : code == 4 ? -d.y : -d.z;
  inside = matched && dn > 0.f;  // dot(normal, d) > 0
  return (smallest_tmax > largest_tmin) && (dist > 0);
}
// ---- fused boundary event ----------------------------------------------------------------
// The medium box is axis aligned, so the six shading frames of Frame::from_z (CVRMath.h:69-75)
// are signed permutations; written out per face code instead of normalize / cross / dot:
//   code  x-axis    y-axis     z-axis (= normal)
//   0 +x  (0,1,0)   (0,0,1)    (1,0,0)        3 -x  (0,1,0)  (0,0,-1)  (-1,0,0)
//   1 +y  (1,0,0)   (0,0,-1)   (0,1,0)        4 -y  (1,0,0)  (0,0,1)   (0,-1,0)
//   2 +z  (1,0,0)   (0,1,0)    (0,0,1)        5 -z  (1,0,0)  (0,-1,0)  (0,0,-1)
CVR_DEV V3 frame_to_local(int c, V3 v) {
  const float lx = (c == 0 || c == 3) ? v.y : v.x;
  const float ly = c == 0 ? v.z : c == 1 ? -v.z : c == 2 ? v.y : c == 3 ? -v.z : c == 4 ? v.z : -v.y;
  const float lz = c == 0 ? v.x : c == 1 ? v.y : c == 2 ? v.z : c == 3 ? -v.x : c == 4 ? -v.y : -v.z;
  return v3(lx, ly, lz);
}
CVR_DEV V3 frame_to_world(int c, V3 a) {
  const float wx = c == 0 ? a.z : c == 3 ? -a.z : a.x;
  const float wy = c == 0 ? a.x : c == 1 ? a.z : c == 2 ? a.y : c == 3 ? a.x : c == 4 ? -a.z : -a.y;
  const float wz = c == 0 ? a.y : c == 1 ? -a.y : c == 2 ? a.z : c == 3 ? -a.y : c == 4 ? a.y : -a.z;
  return v3(wx, wy, wz);
}
// mitsuba_GGX_sampleVNDF + sampleVisible11 (GGX.h:85-181) without the angle round trip: the
// reference takes theta = acos(w.z), phi = atan2(w.y, w.x) and then only ever uses tan(theta),
// sin(phi) and cos(phi) -- which are sqrt(1 - z^2) / z, y / len and x / len.  Same draws, same
// branches (theta < 1e-4 can only be the theta = 0 case: acos(0.999999) = 1.4e-3).
CVR_DEV V3 ggx_sample_vndf_fast(V3 wi_in, float ax, float ay, float u1, float u2) {
  const V3 wi = normalize(v3(ax * wi_in.x, ay * wi_in.y, wi_in.z));
  float sin_phi = 0.f, cos_phi = 1.f, sx, sy;
  if (wi.z < 0.999999f) {
    const float inv_len = rsqrtf(wi.x * wi.x + wi.y * wi.y);
    cos_phi = wi.x * inv_len, sin_phi = wi.y * inv_len;
    const float tan_t = __fdividef(sqrtf(fmaxf(0.f, 1.f - wi.z * wi.z)), wi.z), cot_t = __fdividef(1.0f, tan_t);
    const float G1 = __fdividef(2.0f, 1.0f + sqrtf(1.0f + tan_t * tan_t));
    float A = __fdividef(2.0f * u1, G1) - 1.0f;
    if (fabsf(A) == 1) A -= copysignf(1.0f, A) * CVR_EPS;
    const float tmp = __fdividef(1.0f, A * A - 1.0f);
    const float D = sqrtf(fmaxf(0.0f, (tan_t * tan_t * tmp * tmp) - ((A * A - tan_t * tan_t) * tmp)));
    const float s1 = (tan_t * tmp) - D, s2 = (tan_t * tmp) + D;
    sx = (A < 0.0f || s2 > cot_t) ? s1 : s2;
    float S, v = u2;
    if (v > 0.5f) {
      S = 1.0f;
      v = 2.0f * (v - 0.5f);
    } else {
      S = -1.0f;
      v = 2.0f * (0.5f - v);
    }
    const float z = __fdividef(v * (v * (v * (-0.365728915865723f) + 0.790235037209296f) - 0.424965825137544f) + 0.000152998850436920f,
                               v * (v * (v * (v * 0.169507819808272f - 0.397203533833404f) - 0.232500544458471f) + 1.0f) - 0.539825872510702f);
    sy = S * z * sqrtf(1.0f + (sx * sx));
  } else {  // normal incidence: the theta < 1e-4 branch (GGX.h:94-100)
    const float r = sqrtf(fmaxf(0.0f, __fdividef(u1, 1.0f - u1)));
    float sp, cp;
    __sincosf(2.0f * CVR_PI * u2, &sp, &cp);
    sx = r * cp, sy = r * sp;
  }
  float rx = ((cos_phi * sx) - (sin_phi * sy)) * ax, ry = ((sin_phi * sx) + (cos_phi * sy)) * ay;
  const float nrm = rsqrtf((rx * rx) + (ry * ry) + 1.0f);
  return v3(-rx * nrm, -ry * nrm, nrm);
}
// GGX_sample (GGX.h:265-326) on top of it; Fresnel, reflect / refract and G1 keep their form
// (cvr_device.cuh).  `wo` aliases the ray direction like in the reference.
template <class RNG>
CVR_DEV bool ggx_sample_fast(float ax, float ay, float eta, V3 wi, RNG& rng, V3& wo, float& weight) {
  if (wi.z == 0.f) {
    weight = 0;
    return false;
  }
  weight = 1.0f;
  const float sign = wi.z > 0.f ? 1.0f : -1.0f;
  const float u1 = rng.next();
  const float u2 = rng.next();
  const V3 wh = ggx_sample_vndf_fast(sign * wi, ax, ay, u1, u2);
  float whdotwt = 0.f;
  const float whdotwi = dot(wh, wi);
  const float F = fresnel_dielectric(eta, whdotwi, whdotwt);
  if (rng.next() <= F) {
    const float c2 = 2.f * whdotwi;
    wo = v3(c2 * wh.x - wi.x, c2 * wh.y - wi.y, c2 * wh.z - wi.z);
    if (wi.z * wo.z <= 0) {
      weight = 0.0f;
      return false;
    }
  } else {
    if (whdotwt == 0.0f) {
      weight = 0.0f;
      return false;
    }
    float e = eta;
    if (whdotwt < 0) e = __fdividef(1.0f, e);
    wo = wh * (whdotwi * e + whdotwt) - wi * e;
    if (wi.z * wo.z >= 0) {
      weight = 0.0f;
      return false;
    }
  }
  weight *= ggx_g1(ax, ay, wo, wh);
  return true;
}

// ---- cold code out of line ----------------------------------------------------------------------
// The fused kernels are ~54 KB of SASS against a 32 KB L1.5 instruction cache (and ~6 KB of L0 per scheduler), and 28
// warps per SM sit in different parts of it: ncu shows 0.3 (hetvol) ... 1.2 (manix) warps per issue stalled on
// `no_instruction`.  Code that never runs in the reference's configurations is therefore kept OUT of the instruction
// stream of the events (`__noinline__`: a call instead of 229 / 475 inlined instructions).  Measured on B200 (1024^2 x
// 32 spp, kernel ms, gpurun calls AC / AD, profiles/r2_code_layout_ab.txt):
//   anisotropic Henyey-Greenstein (g != 0; the reference's media are g == 0, Q5) out of line, every kernel:
//     bucky 5.75 -> 5.42, hetvol 26.75 -> 26.33, manix 11.87 -> 11.36, fBm 512^3 20.65 -> 19.79, fBm 1024^3 20.72 -> 19.89,
//     sparse 1024^3 12.94 -> 12.78 (-1 ... -6 %);
//   the counter flush at kernel exit (70 shuffles) out of line as well: manix -> 10.92, fBm -> 19.28 / 19.48, sparse ->
//     12.50 (-3 ... -8 % in total) but hetvol -> 27.23 (+2 %): only in the skip-table kernels (the ones that run volumes
//     beyond the L2).  Besides the code moved, the call changes where the lane counters live: their address escapes, so
//     they stay in LOCAL memory between rounds (52 LDL / STL in the kernel, none inside the pair loop: ptxas loads them
//     into registers around it), which frees registers in the event code.  The skip-table kernels have the L1 data pipe to
//     spare for that (~50 % busy); hetvol's kernel does not (84 %).
// No arithmetic changes: every path stays bit-identical.  What does NOT work: an event that takes the path registers by
// reference out of line (do_boundary, start_path: the whole PathRegs then lives in local memory, +18 ... +79 %).
#ifndef CVR_COLD_NOINLINE
#define CVR_COLD_NOINLINE 1
#endif
#if CVR_COLD_NOINLINE
__device__ __noinline__ V3 hg_sample_cold(V3 dir, float g, float e1, float e2) { return hg_sample(dir, g, e1, e2); }
#else
CVR_DEV V3 hg_sample_cold(V3 dir, float g, float e1, float e2) { return hg_sample(dir, g, e1, e2); }
#endif
CVR_DEV V3 hg_sample_fast(V3 dir, float g, float e1, float e2) {
  if (fabsf(g) > CVR_EPS) return hg_sample_cold(dir, g, e1, e2);  // anisotropic phase: exact path
  float cos_theta = 1.0f - 2.0f * e1;
  float sin_theta = sqrtf(fmaxf(0.0f, 1.0f - cos_theta * cos_theta));
  float sp, cp;
  __sincosf(CVR_TWOPI * e2, &sp, &cp);
  float inv_norm = rsqrtf(dir.x * dir.x + dir.z * dir.z);
  V3 v1 = v3(dir.z * inv_norm, 0.0f, -dir.x * inv_norm);
  V3 v2 = cross(dir, v1);
  return sin_theta * cp * v1 + sin_theta * sp * v2 + cos_theta * dir;
}

// ---- per-path event log (parity hook, cvr_trace_paths with a log buffer) -------------------
// One entry per loop iteration of the path that ended in a scatter or boundary event, plus the
// final escape: (code, XORWOW draw counter `d` when the event starts), the same record the CPU oracle
// of the test suite keeps, so that a test can show where a path whose radiance differs from the CPU oracle's
// left the common event prefix.
enum : uint32_t { EV_SCATTER = 1, EV_BOUNDARY = 2, EV_ESCAPE = 3, EVF_OK = 16, EVF_WO_NEG = 32, EVF_WI_NEG = 64, EVF_KILLED = 128,
                  EVF_ZERO = 256 /* the throughput is exactly zero after the event (roulette then always ends the path) */ };
CVR_DEV uint32_t draw_tag(const Xorwow& g) { return g.d; }
CVR_DEV uint32_t draw_tag(const Philox& g) { return g.c0 * 4u - g.have; }
CVR_DEV uint32_t draw_tag(const PhiloxCB& g) { return g.ctr * 4u - g.have; }
CVR_DEV void log_path_event(const KernelParams& P, uint32_t path_lo, uint32_t index, uint32_t code, uint32_t tag) {
  if (index < P.log_cap) P.path_log[(size_t)path_lo * P.log_cap + index] = make_uint2(code, tag);
}

// ---- regeneration: camera ray for launch-global work item g (A1/A2 prologue) ----
template <int RNGM, class Rng>
CVR_DEV void start_path(const KernelParams& P, unsigned long long g, unsigned long long per_tile,
                        PathRegs<Rng>& R) {
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
  if (RNGM == RNG_XORWOW_PATH) R.rng.init((int32_t)(seed_k + (uint32_t)path_id));
  if (RNGM == RNG_PHILOX) R.rng.init((unsigned long long)seed_k + path_id);
  uint32_t image_id = (uint32_t)(path_id % P.npix);
  if (P.regen_block) {
    // regen_order=block (experiment, section 3.6 of DESIGN.md): consecutive path ids walk 8 x 4 pixel blocks instead of
    // rows, so the 32 paths a warp regenerates at once start on a compact patch of the image -- the 2-D analogue of the
    // reference's Morton-ordered regeneration.  A bijection on the tile's pixels (tile_w % 8 == 0, tile_h % 4 == 0 checked
    // by the host); which pixel a stream lands on changes, so parity of this mode is statistical.
    const uint32_t bw = P.tile_w >> 3, blk = image_id >> 5, in = image_id & 31u;
    image_id = ((blk / bw) * 4u + (in >> 3)) * P.tile_w + (blk % bw) * 8u + (in & 7u);
  }
  float u0 = R.rng.next();
  float u1 = R.rng.next();
  camera_ray(P.cam, image_id, off_x, off_y, u0, u1, R.o, R.d);
  R.thr_x = R.thr_y = R.thr_z = 1.f;
  R.bounces = 0;
  R.path_lo = (uint32_t)g;
  if (P.out_full) {
    uint32_t px = image_id % P.tile_w, py = image_id / P.tile_w;
    R.out_idx = (py + off_y) * P.out_stride + (px + off_x);
  } else {
    R.out_idx = image_id;
  }
  R.state = S_ISECT;
}

// ---- intersect (A5) + escape accumulation (A13) ----
template <bool COUNT, bool FAST = false, bool LOG = false, class Rng>
CVR_DEV void do_isect(const KernelParams& P, PathRegs<Rng>& R, LaneCounters& C) {
  if (COUNT) ++C.bounces;
  bool inside, hit;
  int ncode;
  if (FAST) {
    hit = box_intersect_code(P.med.box_min, P.med.box_max, R.o, R.d, R.dist, ncode, inside);
  } else {
    V3 normal = v3(0, 0, 0);
    hit = box_intersect(P.med.box_min, P.med.box_max, R.o, R.d, R.dist, normal, inside);
    ncode = code_from_normal(normal);
  }
  if (!hit) {
    // escaped: throughput * Le, Le == 1 (Medium.h:174-177)
    float rx = R.thr_x * 1.f, ry = R.thr_y * 1.f, rz = R.thr_z * 1.f;
    // the reference lets a NaN throughput through (u == 1.0 in sampleVisible11)
    const bool keep = !P.fix_nan || (fabsf(rx) <= 3.0e38f && fabsf(ry) <= 3.0e38f && fabsf(rz) <= 3.0e38f);
    if (P.per_path) P.per_path[R.path_lo] = make_float4(rx, ry, rz, 1.f);
    if (LOG && P.path_log) log_path_event(P, R.path_lo, R.bounces, EV_ESCAPE, draw_tag(R.rng));
    if (P.out && keep) {
      float4* px = P.out + R.out_idx;
      atomicAdd(&px->x, rx);
      atomicAdd(&px->y, ry);
      atomicAdd(&px->z, rz);
      px->w = 1.f;
    }
    if (COUNT) ++C.esc;
    if (P.rr_after_escape && P.rr) (void)R.rng.next();
    R.state = S_IDLE;
  } else if (!inside) {
    R.ncode = ncode;
    R.state = S_BOUNDARY;
  } else {
    R.ncode = ncode;
    R.t = 0.f;
    R.state = S_TRACK;
  }
}

// ---- one Woodcock step (A6, A7): Utilities.cuh:146-152 ----
template <int LAYOUT, bool COUNT, bool FAST = false, class Rng>
CVR_DEV void do_track_step(const KernelParams& P, const TrackInv& I, PathRegs<Rng>& R, LaneCounters& C) {
  if (FAST && LAYOUT != LAYOUT_LINEAR) {
    float u = R.rng.next();
    R.t = fmaf(lg2_fast(fmaxf(u, CVR_EPS)), I.neg_ln2_inv_sigmat, R.t);
    V3 p = v3(fmaf(R.t, R.d.x, R.o.x), fmaf(R.t, R.d.y, R.o.y), fmaf(R.t, R.d.z, R.o.z));
    float dens = density_at_fast<LAYOUT>(P.med, I, p);
    if (COUNT) ++C.dens;
    bool go_on = (R.t <= R.dist);
    if (go_on) go_on = (dens * I.sig_ratio < R.rng.next());
    if (!go_on) R.state = (R.t < R.dist) ? S_SCATTER : S_BOUNDARY;
    return;
  }
  float u = R.rng.next();
  R.t += -logf(fmaxf(u, CVR_EPS)) * I.inv_max_sigmat;
  V3 coord = (R.o + (R.t * R.d)) - v3(I.qx, I.qy, I.qz);
  float event_density = P.med.scale * density_at<LAYOUT>(P.med, I, coord);
  if (COUNT) ++C.dens;
  bool go_on = (R.t <= R.dist);
  if (go_on) go_on = (event_density * I.inv_max_sigmat < R.rng.next());
  if (!go_on) R.state = (R.t < R.dist) ? S_SCATTER : S_BOUNDARY;
}

// ---- the Woodcock step of the fast queued kernel -----------------------------------------
// Same draws in the same order as do_track_step, restructured for issue slots:
//  * the ray is carried in GRID space (g0 + t*gd, set up once per batch), so a step needs 3
//    FFMA for the lookup coordinate instead of 6 and no constant-bank operands;
//  * both uniforms of the step are drawn up front (straight-line code, no divergent branch
//    around the second XORWOW update).  The reference does not consume the second uniform
//    when the step left the medium (Utilities.cuh:148-152: the && short-circuits); in that
//    case it is PARKED in R.t for the boundary event that follows (state S_BOUNDARY_P), which
//    takes it as its first draw -- the path sees exactly the reference's sequence.
struct GridRay {
  float g0x, g0y, g0z, gdx, gdy, gdz;
};
CVR_DEV GridRay grid_ray(const TrackInv& I, const V3& o, const V3& d) {
  GridRay G;
  G.g0x = fmaf(o.x, I.rx, I.nqrx), G.g0y = fmaf(o.y, I.ry, I.nqry), G.g0z = fmaf(o.z, I.rz, I.nqrz);
  G.gdx = d.x * I.rx, G.gdy = d.y * I.ry, G.gdz = d.z * I.rz;
  // opaque to the optimiser: otherwise ptxas keeps o/d live and re-derives these six values
  // inside the loop (6 extra instructions + constant loads per step)
  asm volatile("" : "+f"(G.g0x), "+f"(G.g0y), "+f"(G.g0z), "+f"(G.gdx), "+f"(G.gdy), "+f"(G.gdz));
  return G;
}
template <int LAYOUT, bool COUNT, bool SKIP = false>
CVR_DEV void track_step_fast(const KernelParams& P, const TrackInv& I, const GridRay& G, PathRegs<Xorwow>& R,
                             LaneCounters& C, const SkipTab& S = SkipTab()) {
  const float u = R.rng.next();
  const uint32_t r2 = R.rng.next_u32();
  const float u2 = r2 * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
  R.t = fmaf(lg2_fast(fmaxf(u, CVR_EPS)), I.neg_ln2_inv_sigmat, R.t);
  CellFetch F;
  const bool skipped = cell_fetch_skip<LAYOUT, SKIP>(P, I, S, fmaf(R.t, G.gdx, G.g0x), fmaf(R.t, G.gdy, G.g0y),
                                                     fmaf(R.t, G.gdz, G.g0z), r2, F);
  const float dens = trilerp_fast(F.v, F.fx, F.fy, F.fz);
  if (COUNT) ++C.dens;
  if (COUNT && SKIP) C.skip += skipped ? 1u : 0u;
  // select form (no divergent branch): inside ? (accepted ? event : keep tracking) : parked boundary
  const bool inside = R.t <= R.dist;
  const bool accepted = !skipped && !(dens * I.sig_ratio < u2);
  const int ev = (R.t < R.dist) ? S_SCATTER : S_BOUNDARY;
  R.state = inside ? (accepted ? ev : S_TRACK) : S_BOUNDARY_P;
  R.t = inside ? R.t : u2;
}

// Two Woodcock steps at once, the second one SPECULATIVE: the position of step 2 depends
// only on the uniforms (not on the density of step 1), so both 256-bit cell loads are
// issued back to back and a lane has two L2 round trips in flight (the loop is bound by
// the load latency of a dependent chain, not by issue slots -- DESIGN.md 3.1).  If step 1
// ended the segment (collision or exit) the second step never happened: its lookup is
// counted as speculative and the generator is rolled back by two draws.  The state after 2
// draws (v2,v3,v4,n1,n2) shares three words with the state after 4 (v4,n1,n2,n3,n4), so
// the roll-back is six selects on two saved words, not an inverse computation.
template <int LAYOUT, bool COUNT, bool SKIP = false>
CVR_DEV void track_pair_fast(const KernelParams& P, const TrackInv& I, const GridRay& G, PathRegs<Xorwow>& R,
                             LaneCounters& C, const SkipTab& S = SkipTab()) {
  Xorwow& g = R.rng;
  const uint32_t s2 = g.v2, s3 = g.v3;
  const float u1 = g.next();
  const uint32_t r1 = g.next_u32();
  const float u2 = g.next();
  const uint32_t r2 = g.next_u32();
  // curand_uniform of the two accept words (Xorwow::next)
  const float w1 = r1 * 2.3283064e-10f + (2.3283064e-10f / 2.0f), w2 = r2 * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
  const float t1 = fmaf(lg2_fast(fmaxf(u1, CVR_EPS)), I.neg_ln2_inv_sigmat, R.t);
  const float t2 = fmaf(lg2_fast(fmaxf(u2, CVR_EPS)), I.neg_ln2_inv_sigmat, t1);
  CellFetch F1, F2;
  const bool k1 = cell_fetch_skip<LAYOUT, SKIP>(P, I, S, fmaf(t1, G.gdx, G.g0x), fmaf(t1, G.gdy, G.g0y),
                                                fmaf(t1, G.gdz, G.g0z), r1, F1);
  const bool k2 = cell_fetch_skip<LAYOUT, SKIP>(P, I, S, fmaf(t2, G.gdx, G.g0x), fmaf(t2, G.gdy, G.g0y),
                                                fmaf(t2, G.gdz, G.g0z), r2, F2);
  // both loads must be in flight before the first blend: tie the two results together so
  // that ptxas cannot consume load 1 (and reuse its registers) before load 2 is issued
  // (a REAL data dependency: an empty asm leaves nothing for ptxas to order; density values
  // are finite, so 0 * v is exactly 0 and the blend of step 1 is unchanged)
  F1.v[0] = __fmaf_rn(0.0f, F2.v[7], F1.v[0]);
  F1.v[4] = __fmaf_rn(0.0f, F2.v[7], F1.v[4]);
  const float dens1 = trilerp_fast(F1.v, F1.fx, F1.fy, F1.fz);
  const float dens2 = trilerp_fast(F2.v, F2.fx, F2.fy, F2.fz);
  const bool in1 = t1 <= R.dist, in2 = t2 <= R.dist;
  // a skipped step is a certain null collision (SkipTab): its blend is of an unrelated cell
  const bool acc1 = !k1 && !(dens1 * I.sig_ratio < w1), acc2 = !k2 && !(dens2 * I.sig_ratio < w2);
  const bool cont1 = in1 && !acc1;  // step 1 was a null collision inside the medium
  // the step that decides this lane's fate: step 2 if step 1 continued, else step 1.  Its `inside`
  // and `accepted` flags as pure predicate logic (a ?: between bools costs ptxas a SEL / LOP3 /
  // ISETP round trip through an integer register each):
  //   ind  = cont1 ? in2  : in1   = in1 && (acc1 || in2)
  //   accd = cont1 ? acc2 : acc1  = acc1 || (cont1 && acc2)
  const float td = cont1 ? t2 : t1, wd = cont1 ? w2 : w1;
  const bool ind = in1 & (acc1 | in2), accd = acc1 | (cont1 & acc2);
  int st = (td < R.dist) ? S_SCATTER : S_BOUNDARY;
  st = accd ? st : S_TRACK;
  R.state = ind ? st : S_BOUNDARY_P;
  R.t = ind ? td : wd;
  // roll the generator back by two draws when step 2 never happened
  g.v4 = cont1 ? g.v4 : g.v2;
  g.v3 = cont1 ? g.v3 : g.v1;
  g.v2 = cont1 ? g.v2 : g.v0;
  g.v1 = cont1 ? g.v1 : s3;
  g.v0 = cont1 ? g.v0 : s2;
  g.d = cont1 ? g.d : g.d - 2u * 362437u;
  if (COUNT) {
    C.pairs += 1u;
    C.cont += cont1 ? 1u : 0u;
    if (SKIP) C.skip += (k1 ? 1u : 0u) + (k2 ? 1u : 0u);  // cell loads not issued (of the 2 per pair, speculative one included)
  }
}

// The same pair of Woodcock steps on the counter-based stream (rng=philox): ONE Philox block is the
// four words of the pair (step 1: flight, accept; step 2: flight, accept).  When step 1 ends the
// segment the two words of step 2 are simply dropped -- the next consumer takes a fresh block --
// so there is no generator roll-back, no parked uniform (a path that left the medium goes to
// S_BOUNDARY and the boundary event draws its own block) and nothing but the block counter to
// write back.  Statistical parity only (the draw ORDER differs from the reference's).
template <int LAYOUT, bool COUNT, bool SKIP = false>
CVR_DEV void track_pair_cb(const KernelParams& P, const TrackInv& I, const GridRay& G, PathRegs<PhiloxCB>& R,
                           LaneCounters& C, const SkipTab& S = SkipTab()) {
  uint32_t q1, r1, q2, r2;
  R.rng.block(q1, r1, q2, r2);
  const float k = 2.3283064e-10f, h = 2.3283064e-10f / 2.0f;
  const float u1 = q1 * k + h, w1 = r1 * k + h, u2 = q2 * k + h, w2 = r2 * k + h;
  const float t1 = fmaf(lg2_fast(fmaxf(u1, CVR_EPS)), I.neg_ln2_inv_sigmat, R.t);
  const float t2 = fmaf(lg2_fast(fmaxf(u2, CVR_EPS)), I.neg_ln2_inv_sigmat, t1);
  CellFetch F1, F2;
  const bool k1 = cell_fetch_skip<LAYOUT, SKIP>(P, I, S, fmaf(t1, G.gdx, G.g0x), fmaf(t1, G.gdy, G.g0y),
                                                fmaf(t1, G.gdz, G.g0z), r1, F1);
  const bool k2 = cell_fetch_skip<LAYOUT, SKIP>(P, I, S, fmaf(t2, G.gdx, G.g0x), fmaf(t2, G.gdy, G.g0y),
                                                fmaf(t2, G.gdz, G.g0z), r2, F2);
  F1.v[0] = __fmaf_rn(0.0f, F2.v[7], F1.v[0]);  // both loads in flight before the first blend (track_pair_fast)
  F1.v[4] = __fmaf_rn(0.0f, F2.v[7], F1.v[4]);
  const float dens1 = trilerp_fast(F1.v, F1.fx, F1.fy, F1.fz);
  const float dens2 = trilerp_fast(F2.v, F2.fx, F2.fy, F2.fz);
  const bool in1 = t1 <= R.dist, in2 = t2 <= R.dist;
  const bool acc1 = !k1 && !(dens1 * I.sig_ratio < w1), acc2 = !k2 && !(dens2 * I.sig_ratio < w2);
  const bool cont1 = in1 && !acc1;
  const float td = cont1 ? t2 : t1;
  const bool ind = in1 & (acc1 | in2), accd = acc1 | (cont1 & acc2);
  int st = (td < R.dist) ? S_SCATTER : S_BOUNDARY;
  st = accd ? st : S_TRACK;
  R.state = ind ? st : S_BOUNDARY;
  R.t = td;
  if (COUNT) {
    C.pairs += 1u;
    C.cont += cont1 ? 1u : 0u;
    if (SKIP) C.skip += (k1 ? 1u : 0u) + (k2 ? 1u : 0u);
  }
}

// ---- one step of LOCAL-majorant delta tracking ("tracking=local") ----------------------
// Same unbiased estimator as Woodcock tracking, but the majorant is piecewise constant
// over bricks of CVR_BRICK^3 lookup cells (the "majorant mip" of the design): inside a
// brick with majorant mu the free-flight step is drawn with sigma = scale*mu; a step that
// leaves the brick moves the path to the brick face without a lookup (memoryless
// restart); bricks with mu == 0 are skipped without a draw.  Fewer null collisions =>
// fewer lookups per path.  RNG consumption differs from the reference's global-majorant
// loop, so parity of this mode is statistical.
//
// BrickWalk is the per-lane state of the walk through the brick grid.  A brick is located
// FROM SCRATCH (probe point, floors, three divides, both mip levels) at the start of a batch,
// when the walk enters another 64^3 super-brick, and while it crosses empty super-bricks in
// one stride; inside a non-empty super-brick the walk is an INCREMENTAL 3-D DDA: the axis
// whose exit parameter is the smallest steps by one brick, only that axis' exit parameter is
// recomputed (one FMA with the cached reciprocal), and one majorant is fetched.
struct BrickWalk {
  float texit, mu;      // exit parameter and majorant of the brick the path is in
  float jx, jy, jz;     // fine brick coordinates (unclamped, grid space)
  float ex, ey, ez;     // exit parameter per axis
  float rx, ry, rz;     // 1 / gd per axis (3e38 where the ray is parallel to the faces)
  int fine;             // 1: (jx,jy,jz) / (ex,ey,ez) describe a fine brick of a non-empty super-brick
};
CVR_DEV void brick_walk_reset(BrickWalk& W) {
  W.texit = -1.0f, W.mu = 0.f, W.fine = 0;
  W.jx = W.jy = W.jz = W.ex = W.ey = W.ez = W.rx = W.ry = W.rz = 0.f;
}
// majorant of fine brick (jx,jy,jz): bricks outside the grid use the far-edge brick, like the
// lookup's clamping (Q2: both out-of-range sides read the far edge)
CVR_DEV float fine_majorant(const TrackInv& I, float jx, float jy, float jz) {
  const uint32_t lx = I.nx >> CVR_BRICK_LOG2, ly = I.ny >> CVR_BRICK_LOG2, lz = I.nz >> CVR_BRICK_LOG2;
  const uint32_t bx = min((uint32_t)(int)jx, lx), by = min((uint32_t)(int)jy, ly), bz = min((uint32_t)(int)jz, lz);
  return __ldg(I.majorant + bx + I.mx * (by + I.my * bz));
}

template <int LAYOUT, bool COUNT, class Rng>
CVR_DEV void do_track_step_local(const KernelParams& P, const TrackInv& I, const GridRay& G, PathRegs<Rng>& R,
                                 LaneCounters& C, BrickWalk& W) {
  // (Measured and rejected: hopping over up to 8 empty bricks inside one call instead of one
  // per iteration of the warp's step loop -- the hopping lanes hold the warp up: hetvol 954 ->
  // 812, manix 2362 -> 2013, sparse 2048^3 1076 -> 620 Msamples/s.)
  if (W.texit <= R.t) {
    bool located = false;
    if (W.fine) {
      // incremental step: leave through the axis with the smallest exit parameter
      const bool ax = W.ex <= W.ey && W.ex <= W.ez, ay = !ax && W.ey <= W.ez;
      const float gd = ax ? G.gdx : ay ? G.gdy : G.gdz;
      const float g0 = ax ? G.g0x : ay ? G.g0y : G.g0z;
      const float r = ax ? W.rx : ay ? W.ry : W.rz;
      float j = ax ? W.jx : ay ? W.jy : W.jz;
      const float s = gd > 0.f ? 1.0f : -1.0f;
      j += s;
      // still inside the same 64^3 super-brick?  (entering brick 8m from below or 8m+7 from above leaves it)
      const int jm = (int)j & 7;
      located = (s > 0.f) ? (jm != 0) : (jm != 7);
      if (located) {
        const float plane = (s > 0.f ? j + 1.0f : j) * CVR_BRICK - 1.0f;
        const float e = (plane - g0) * r;
        if (ax)
          W.jx = j, W.ex = e;
        else if (ay)
          W.jy = j, W.ey = e;
        else
          W.jz = j, W.ez = e;
        W.texit = fmaxf(fminf(fminf(W.ex, W.ey), W.ez), W.texit);
        W.mu = fine_majorant(I, W.jx, W.jy, W.jz);
      }
    }
    if (!located) {
      // from scratch.  grid-space ray g(t) = g0 + t gd; brick faces sit at g = CVR_BRICK*j - 1
      const float tp = R.t + 1e-6f + 1e-6f * fabsf(R.t);  // probe just inside the next brick
      const float gx = fmaf(tp, G.gdx, G.g0x), gy = fmaf(tp, G.gdy, G.g0y), gz = fmaf(tp, G.gdz, G.g0z);
      // lookup cell of the probe point, clamped like the lookup itself
      const uint32_t kx = min((uint32_t)((int)floorf(gx) + 1), I.nx), ky = min((uint32_t)((int)floorf(gy) + 1), I.ny),
                     kz = min((uint32_t)((int)floorf(gz) + 1), I.nz);
      // two-level majorant mip: an empty 64^3 super-brick is crossed in one stride, otherwise
      // the 8^3 brick the point is in.  Both levels are fetched together: the walk is a chain of
      // dependent loads, a second round trip per brick would double its latency
      const uint32_t b2 = (kx >> (CVR_BRICK_LOG2 + 3)) +
                          I.m2x * ((ky >> (CVR_BRICK_LOG2 + 3)) + I.m2y * (kz >> (CVR_BRICK_LOG2 + 3)));
      const uint32_t b1 = (kx >> CVR_BRICK_LOG2) + I.mx * ((ky >> CVR_BRICK_LOG2) + I.my * (kz >> CVR_BRICK_LOG2));
      const float mu2 = __ldg(I.majorant2 + b2), mu1 = __ldg(I.majorant + b1);
      const bool coarse_empty = !(mu2 > 0.f);
      const float size = coarse_empty ? (float)(CVR_BRICK * 8) : (float)CVR_BRICK;
      const float inv_b = coarse_empty ? 1.0f / (CVR_BRICK * 8) : 1.0f / CVR_BRICK;
      W.jx = floorf((gx + 1.0f) * inv_b), W.jy = floorf((gy + 1.0f) * inv_b), W.jz = floorf((gz + 1.0f) * inv_b);
      const float big = 3.0e38f;
      W.rx = G.gdx != 0.f ? __fdividef(1.0f, G.gdx) : big;
      W.ry = G.gdy != 0.f ? __fdividef(1.0f, G.gdy) : big;
      W.rz = G.gdz != 0.f ? __fdividef(1.0f, G.gdz) : big;
      // exit parameter of this (super-)brick along the ray
      const float bx = (G.gdx > 0.f ? W.jx + 1.0f : W.jx) * size - 1.0f;
      const float by = (G.gdy > 0.f ? W.jy + 1.0f : W.jy) * size - 1.0f;
      const float bz = (G.gdz > 0.f ? W.jz + 1.0f : W.jz) * size - 1.0f;
      W.ex = G.gdx != 0.f ? (bx - G.g0x) * W.rx : big;
      W.ey = G.gdy != 0.f ? (by - G.g0y) * W.ry : big;
      W.ez = G.gdz != 0.f ? (bz - G.g0z) * W.rz : big;
      W.texit = fmaxf(fminf(fminf(W.ex, W.ey), W.ez), tp);  // always makes progress
      W.mu = coarse_empty ? 0.f : mu1;
      W.fine = coarse_empty ? 0 : 1;
    }
  }
  const bool last = W.texit >= R.dist;  // this brick reaches the end of the segment
  const float tend = last ? R.dist : W.texit;
  if (W.mu <= 0.f) {  // empty brick: skip it
    R.t = W.texit;
    if (last) R.state = S_BOUNDARY;
    return;
  }
  float u = R.rng.next();
  float tn = fmaf(lg2_fast(fmaxf(u, CVR_EPS)), __fdividef(-0.69314718055994530942f, P.med.scale * W.mu), R.t);
  if (tn >= tend) {  // left the brick (or the medium) without a collision
    R.t = W.texit;
    if (last) R.state = S_BOUNDARY;
    return;
  }
  R.t = tn;
  float dens = density_at_grid<LAYOUT>(P.med, I, fmaf(R.t, G.gdx, G.g0x), fmaf(R.t, G.gdy, G.g0y), fmaf(R.t, G.gdz, G.g0z));
  if (COUNT) ++C.dens;
  if (dens >= R.rng.next() * W.mu) R.state = S_SCATTER;  // real collision with probability dens / mu
}

// ---- Russian roulette (NaiveVolPTsk_kernel.cuh:75-84) + bounce cap ----
template <bool FAST = false, class Rng, class Draw>
CVR_DEV void do_roulette(const KernelParams& P, PathRegs<Rng>& R, Draw& rng) {
  R.state = S_ISECT;
  if (P.rr) {
    float p_survive = fminf(1.f, fmaxf(fmaxf(R.thr_x, R.thr_y), R.thr_z));
    if (rng.next() > p_survive) R.state = S_IDLE;
    if (FAST) {
      float ip = __fdividef(1.f, p_survive);
      R.thr_x *= ip, R.thr_y *= ip, R.thr_z *= ip;
    } else {
      R.thr_x = R.thr_x * 1.f / p_survive;
      R.thr_y = R.thr_y * 1.f / p_survive;
      R.thr_z = R.thr_z * 1.f / p_survive;
    }
  }
  ++R.bounces;
  if (P.max_bounces && R.bounces >= P.max_bounces) R.state = S_IDLE;
}

// ---- scatter event (A8, A9): NaiveVolPTsk_kernel.cuh:67-71 / Regeneration...:212-216 ----
template <int LAYOUT, bool COUNT, bool FAST = false, bool LOG = false, class Rng>
CVR_DEV void do_scatter(const KernelParams& P, PathRegs<Rng>& R, LaneCounters& C) {
  const uint32_t ev_tag = LOG ? draw_tag(R.rng) : 0u, ev_index = R.bounces;
  if (P.pullback)
    R.o = R.o + R.d * R.t - R.d * CVR_EPS;
  else
    R.o = R.o + R.d * R.t;
  V3 albedo;
  if (FAST && LAYOUT != LAYOUT_LINEAR) {
    V3 ac = v3((R.o.x - P.med.box_min.x) * P.inv.aix, (R.o.y - P.med.box_min.y) * P.inv.aiy,
               (R.o.z - P.med.box_min.z) * P.inv.aiz);
    albedo = P.med.albedo_const ? v3(P.med.albedo_r, P.med.albedo_g, P.med.albedo_b) : albedo_cell8_fast(P.med, ac);
  } else {
    V3 ac = (R.o - P.med.box_min) / (P.med.box_max - P.med.box_min);
    albedo = albedo_lookup<LAYOUT>(P.med, ac);
  }
  if (COUNT) ++C.alb;
  R.thr_x = R.thr_x * albedo.x, R.thr_y = R.thr_y * albedo.y, R.thr_z = R.thr_z * albedo.z;
  float e1 = R.rng.next();
  float e2 = R.rng.next();
  R.d = FAST ? hg_sample_fast(R.d, P.med.hg_g, e1, e2) : hg_sample(R.d, P.med.hg_g, e1, e2);
  const uint32_t ev_zero = LOG && fmaxf(fmaxf(R.thr_x, R.thr_y), R.thr_z) == 0.f ? EVF_ZERO : 0u;
  do_roulette<FAST>(P, R, R.rng);
  if (LOG && P.path_log)
    log_path_event(P, R.path_lo, ev_index, EV_SCATTER | ev_zero | (R.state == S_IDLE && P.rr ? EVF_KILLED : 0u), ev_tag);
}

// ---- boundary event (A10): NaiveVolPTsk_kernel.cuh:50-65 ----
template <bool FAST = false, bool LOG = false, bool SMALLTRIG = false, class Rng>
CVR_DEV void do_boundary(const KernelParams& P, PathRegs<Rng>& R) {
  // S_BOUNDARY_P: the first uniform of this event was drawn by the tracking loop (parked in t)
  StashRng<Rng> rng{R.rng, R.t, R.state == S_BOUNDARY_P};
  // the parked uniform was drawn ahead of the event: it does not count as consumed yet
  const uint32_t ev_tag = LOG ? draw_tag(R.rng) - (rng.has ? 362437u : 0u) : 0u, ev_index = R.bounces;
  uint32_t ev_code = EV_BOUNDARY;
  float weight = 1;
  if (FAST) {
    const V3 dir = frame_to_local(R.ncode, normalize(v3(-R.d.x, -R.d.y, -R.d.z)));
    R.o = v3(fmaf(R.d.x, R.dist, R.o.x), fmaf(R.d.y, R.dist, R.o.y), fmaf(R.d.z, R.dist, R.o.z));
    // the sampler writes the LOCAL direction into the ray even when it then fails
    // CVR_FAST_GGX=1 swaps in the sampler without the angle round trip (ggx_sample_vndf_fast).  It CHANGES the
    // sampler's arithmetic (the reference's acos / atan2 / tan round trip amplifies one ulp of a nearly axial
    // direction to 6e-4 of the angle, so per-path agreement with the reference would drop on boundary-dominated
    // scenes), which is why it stays off: the fused mode keeps the GGX boundary in the reference's form.
    // Measured as a code-size experiment in round 2 (kernel ms at 1024^2 x 32 spp, profiles/r2_code_layout_ab.txt):
    // bucky 5.75 -> 4.93, manix 11.83 -> 10.60, fBm 1024^3 20.71 -> 19.40, hetvol 26.77 -> 26.64; most of that gain is
    // now had without touching the arithmetic ("cold code out of line" above and sin_small / cos_small).
#ifndef CVR_FAST_GGX
#define CVR_FAST_GGX 0
#endif
#if CVR_FAST_GGX
    if (ggx_sample_fast(P.med.alpha_x, P.med.alpha_y, P.med.eta, dir, rng, R.d, weight)) {
#else
    if (ggx_sample<SMALLTRIG>(P.med.alpha_x, P.med.alpha_y, P.med.eta, dir, rng, R.d, weight)) {
#endif
      ev_code |= EVF_OK | (R.d.z < 0.f ? EVF_WO_NEG : 0u);
      R.thr_x *= weight, R.thr_y *= weight, R.thr_z *= weight;
      R.d = frame_to_world(R.ncode, R.d);
      R.o = v3(fmaf(R.d.x, CVR_EPS, R.o.x), fmaf(R.d.y, CVR_EPS, R.o.y), fmaf(R.d.z, CVR_EPS, R.o.z));
    } else {
      ev_code |= R.d.z < 0.f ? EVF_WO_NEG : 0u;
    }
    ev_code |= dir.z < 0.f ? EVF_WI_NEG : 0u;
  } else {
    Frame frame;
    frame.from_z(normal_from_code(R.ncode));
    V3 dir = frame.to_local(normalize(v3(-R.d.x, -R.d.y, -R.d.z)));
    R.o = R.o + R.d * R.dist;
    // the sampler writes the LOCAL direction into the ray even when it then fails
    if (ggx_sample(P.med.alpha_x, P.med.alpha_y, P.med.eta, dir, rng, R.d, weight)) {
      ev_code |= EVF_OK | (R.d.z < 0.f ? EVF_WO_NEG : 0u);
      R.thr_x *= weight, R.thr_y *= weight, R.thr_z *= weight;
      R.d = frame.to_world(R.d);
      R.o = R.o + R.d * CVR_EPS;
    } else {
      ev_code |= R.d.z < 0.f ? EVF_WO_NEG : 0u;
    }
    ev_code |= dir.z < 0.f ? EVF_WI_NEG : 0u;
  }
  if (LOG && fmaxf(fmaxf(R.thr_x, R.thr_y), R.thr_z) == 0.f) ev_code |= EVF_ZERO;
  do_roulette<FAST>(P, R, rng);
  if (LOG && P.path_log) log_path_event(P, R.path_lo, ev_index, ev_code | (R.state == S_IDLE && P.rr ? EVF_KILLED : 0u), ev_tag);
  // nobody drew (grazing hit with roulette off): give the parked uniform back to the stream
  if (rng.has) R.rng.undo();
}

CVR_DEV void flush_counters_body(const KernelParams& P, LaneCounters& C, unsigned lane) {
  const unsigned FULL = 0xffffffffu;
  C.dens += C.pairs + C.cont;
  C.spec += C.pairs - C.cont;
  for (int s = 16; s > 0; s >>= 1) {
    C.paths += __shfl_xor_sync(FULL, C.paths, s);
    C.bounces += __shfl_xor_sync(FULL, C.bounces, s);
    C.dens += __shfl_xor_sync(FULL, C.dens, s);
    C.alb += __shfl_xor_sync(FULL, C.alb, s);
    C.esc += __shfl_xor_sync(FULL, C.esc, s);
    C.spec += __shfl_xor_sync(FULL, C.spec, s);
    C.skip += __shfl_xor_sync(FULL, C.skip, s);
  }
  if (lane == 0) {  // one atomic per warp per counter
    atomicAdd(&P.ctr->paths, (unsigned long long)C.paths);
    atomicAdd(&P.ctr->bounces, (unsigned long long)C.bounces);
    atomicAdd(&P.ctr->density_lookups, (unsigned long long)C.dens);
    atomicAdd(&P.ctr->albedo_lookups, (unsigned long long)C.alb);
    atomicAdd(&P.ctr->escaped, (unsigned long long)C.esc);
    if (C.spec) atomicAdd(&P.ctr->speculative, (unsigned long long)C.spec);
    if (C.skip) atomicAdd(&P.ctr->skipped, (unsigned long long)C.skip);
  }
}
__device__ __noinline__ void flush_counters_outline(const KernelParams& P, LaneCounters& C, unsigned lane) {
  flush_counters_body(P, C, lane);
}
// OUTLINE: a call instead of the inlined body (the skip-table kernels; "cold code out of line" above)
template <bool COUNT, bool OUTLINE = false>
CVR_DEV void flush_counters(const KernelParams& P, LaneCounters& C, unsigned lane) {
  if (!COUNT) return;
  if (OUTLINE && CVR_COLD_NOINLINE)
    flush_counters_outline(P, C, lane);
  else
    flush_counters_body(P, C, lane);
}

// claim path ids for the idle lanes of this warp with one 64-bit atomic
template <int RNGM, bool COUNT, class Rng>
CVR_DEV void warp_regenerate(const KernelParams& P, unsigned idle_mask, unsigned lane, unsigned long long total,
                             unsigned long long per_tile, bool& exhausted, PathRegs<Rng>& R, LaneCounters& C) {
  const unsigned FULL = 0xffffffffu;
  if (!exhausted) {
    int n = __popc(idle_mask);
    int leader = __ffs(idle_mask) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(P.head, (unsigned long long)n);
    base = __shfl_sync(FULL, base, leader);
    if (R.state == S_IDLE) {
      unsigned long long g = base + __popc(idle_mask & ((1u << lane) - 1u));
      if (g < total) {
        start_path<RNGM>(P, g, per_tile, R);
        if (COUNT) ++C.paths;
      } else {
        R.state = S_DONE;
      }
    }
    exhausted = (base + (unsigned long long)n >= total);
  } else if (R.state == S_IDLE) {
    R.state = S_DONE;
  }
}

// =============================================================================
// Scheduler 1: lane-persistent ("while-while"): a lane keeps its path in registers;
// the warp alternates between the Woodcock loop and the event phase.
// =============================================================================
template <int RNGM, int LAYOUT, bool COUNT>
__global__ void __launch_bounds__(CVR_BLOCK, CVR_MIN_BLOCKS)
    k_volpt(const __grid_constant__ KernelParams P) {
  typedef typename RngSel<RNGM>::type Rng;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;

  PathRegs<Rng> R;
  R.o = v3(0, 0, 0), R.d = v3(0, 0, 1);
  R.thr_x = R.thr_y = R.thr_z = 1.f;
  R.t = R.dist = 0.f;
  R.out_idx = R.path_lo = R.bounces = 0;
  R.ncode = 0;
  R.state = S_IDLE;
  if (RNGM == RNG_XORWOW_THREAD) {
    // RegenerationVolPTsk_kernel.cuh:151,156: Rng rng(seed + tid), once per thread
    uint32_t tid = threadIdx.x + blockDim.x * blockIdx.x;
    R.rng.init((int32_t)(P.seed + tid));
  }
  LaneCounters C;
  const unsigned long long per_tile = P.path_end - P.path_begin;
  const unsigned long long total = per_tile * P.n_launch_tiles;
  const TrackInv& I = P.inv;
  bool exhausted = false;

  for (;;) {
    unsigned idle = __ballot_sync(FULL, R.state == S_IDLE);
    if (idle) warp_regenerate<RNGM, COUNT>(P, idle, lane, total, per_tile, exhausted, R, C);
    if (__all_sync(FULL, R.state == S_DONE)) break;

    if (R.state == S_ISECT) do_isect<COUNT>(P, R, C);

    {
      const int n_parked = __popc(__ballot_sync(FULL, R.state == S_DONE));
      // at least one step per outer iteration, then yield to the event phase as soon
      // as `loop_threshold` lanes are waiting for it
      for (bool first = true;; first = false) {
        unsigned trk = __ballot_sync(FULL, R.state == S_TRACK);
        if (trk == 0) break;
        int waiting = 32 - n_parked - __popc(trk);
        if (!first && waiting >= P.loop_threshold) break;
        if (R.state == S_TRACK) do_track_step<LAYOUT, COUNT>(P, I, R, C);
      }
    }

    if (R.state == S_SCATTER)
      do_scatter<LAYOUT, COUNT>(P, R, C);
    else if (R.state == S_BOUNDARY)
      do_boundary(P, R);
  }
  flush_counters<COUNT>(P, C, lane);
}

// =============================================================================
// Scheduler 2: block-sorted wavefront.  The CVR_BLOCK paths of a CTA live in shared
// memory (80-byte slots).  Every round the slots are counting-sorted by state
// (TRACK, SCATTER, BOUNDARY, IDLE, DONE) so that the 32 lanes of a warp pick up 32
// paths in the SAME state, run that state's event (+ the following intersect), then
// up to `track_steps` Woodcock steps, and store the paths back.  Each path's own
// operation and RNG-draw order is unchanged; only which lane executes it changes.
// =============================================================================
// Field grouping by WHO writes: q0..q2 change only at events ("static"), q3 / q4 after every batch
// ("dynamic": the generator, t and the state).  Every access is then a whole 16-byte vector: the
// first grouping kept t in q0.w and the state in q2.w, so a tracking batch ended with two extra
// scalar stores and began with a scalar and a 64-bit load -- with 80-byte slots a 32-bit access
// of 32 lanes falls on 8 banks only, each about as expensive on the L1 data pipe as a full vector
// (the pipe is the kernel's busiest unit).
struct __align__(16) PathSlot {
  float4 q0;  // o.xyz, dist
  float4 q1;  // d.xyz, out_idx (bits)
  float4 q2;  // thr.xyz, path_lo (bits)
  uint4 q3;   // rng v0..v3
  uint4 q4;   // rng v4, d, t (bits), meta (state | ncode << 3 | bounces << 6)
};

CVR_DEV void slot_store(PathSlot& s, const PathRegs<Xorwow>& R) {
  uint32_t meta = (uint32_t)R.state | ((uint32_t)R.ncode << 3) | (R.bounces << 6);
  s.q0 = make_float4(R.o.x, R.o.y, R.o.z, R.dist);
  s.q1 = make_float4(R.d.x, R.d.y, R.d.z, __uint_as_float(R.out_idx));
  s.q2 = make_float4(R.thr_x, R.thr_y, R.thr_z, __uint_as_float(R.path_lo));
  s.q3 = make_uint4(R.rng.v0, R.rng.v1, R.rng.v2, R.rng.v3);
  s.q4 = make_uint4(R.rng.v4, R.rng.d, __float_as_uint(R.t), meta);
}
CVR_DEV void slot_load(const PathSlot& s, PathRegs<Xorwow>& R) {
  float4 a = s.q0, b = s.q1, c = s.q2;
  uint4 e = s.q3, f = s.q4;
  R.o = v3(a.x, a.y, a.z), R.dist = a.w;
  R.d = v3(b.x, b.y, b.z), R.out_idx = __float_as_uint(b.w);
  R.thr_x = c.x, R.thr_y = c.y, R.thr_z = c.z, R.path_lo = __float_as_uint(c.w);
  R.rng.v0 = e.x, R.rng.v1 = e.y, R.rng.v2 = e.z, R.rng.v3 = e.w;
  R.rng.v4 = f.x, R.rng.d = f.y, R.t = __uint_as_float(f.z);
  const uint32_t meta = f.w;
  R.state = (int)(meta & 7u), R.ncode = (int)((meta >> 3) & 7u), R.bounces = meta >> 6;
}

// The tracking loop only changes t, the state and the generator; everything else is
// written once after the event ("static" part) so that o, d, throughput and the output
// index are dead registers inside the loop, and a pure tracking batch moves 64 + 32 bytes
// per path through shared memory (4 vector loads, 2 vector stores) instead of 80 + 80.
CVR_DEV void slot_load_track(const PathSlot& s, PathRegs<Xorwow>& R) {
  float4 a = s.q0, b = s.q1;
  uint4 e = s.q3, f = s.q4;
  R.o = v3(a.x, a.y, a.z), R.dist = a.w;
  R.d = v3(b.x, b.y, b.z);
  R.rng.v0 = e.x, R.rng.v1 = e.y, R.rng.v2 = e.z, R.rng.v3 = e.w;
  R.rng.v4 = f.x, R.rng.d = f.y, R.t = __uint_as_float(f.z);
  const uint32_t meta = f.w;
  R.state = (int)(meta & 7u), R.ncode = (int)((meta >> 3) & 7u), R.bounces = meta >> 6;
}
CVR_DEV void slot_store_static(PathSlot& s, const PathRegs<Xorwow>& R) {
  s.q0 = make_float4(R.o.x, R.o.y, R.o.z, R.dist);
  s.q1 = make_float4(R.d.x, R.d.y, R.d.z, __uint_as_float(R.out_idx));
  s.q2 = make_float4(R.thr_x, R.thr_y, R.thr_z, __uint_as_float(R.path_lo));
}
CVR_DEV void slot_store_dynamic(PathSlot& s, float t, uint32_t meta, const Xorwow& g) {
  s.q3 = make_uint4(g.v0, g.v1, g.v2, g.v3);
  s.q4 = make_uint4(g.v4, g.d, __float_as_uint(t), meta);
}

// ---- path-slot storage of the warp-private scheduler, per generator ------------------------
// XORWOW: the 80-byte array-of-structs slot above (stride 5 x 16 B: consecutive slots fall on
// different bank groups).  Counter-based stream (rng=philox): 64 bytes per path -- no generator
// state beyond the stream id and the block counter -- as FOUR arrays of 16-byte vectors per warp
// (a 64-byte struct would put every slot of a quarter-warp on two bank groups):
//   q0[W] o.xyz, dist | q1[W] d.xyz, stream lo | q2[W] thr.xyz, out_idx | q3[W] block counter, t, meta, stream hi
// A tracking batch loads q0, q1, q3 and stores q3: 4 vector accesses per path instead of 6.
template <class Rng, int W>
struct WarpSlots;
template <int W>
struct WarpSlots<Xorwow, W> {
  static constexpr size_t kSlotBytes = sizeof(PathSlot);
  PathSlot* s;
  CVR_DEV WarpSlots(unsigned char* warp_base, uint32_t) : s(reinterpret_cast<PathSlot*>(warp_base)) {}
  CVR_DEV void store(unsigned i, const PathRegs<Xorwow>& R) { slot_store(s[i], R); }
  CVR_DEV void load(unsigned i, PathRegs<Xorwow>& R) { slot_load(s[i], R); }
  CVR_DEV void load_track(unsigned i, PathRegs<Xorwow>& R) { slot_load_track(s[i], R); }
  CVR_DEV void store_static(unsigned i, const PathRegs<Xorwow>& R) { slot_store_static(s[i], R); }
  CVR_DEV void store_dynamic(unsigned i, float t, uint32_t meta, const Xorwow& g) { slot_store_dynamic(s[i], t, meta, g); }
};
template <int W>
struct WarpSlots<PhiloxCB, W> {
  static constexpr size_t kSlotBytes = 64;
  float4* q;               // q[k * W + slot]
  uint32_t path_lo_base;   // stream lo - this = index of the path in a single-tile launch (per-path debug output)
  CVR_DEV WarpSlots(unsigned char* warp_base, uint32_t base) : q(reinterpret_cast<float4*>(warp_base)), path_lo_base(base) {}
  CVR_DEV void unpack_dynamic(const float4& e, PathRegs<PhiloxCB>& R) {
    R.rng.ctr = __float_as_uint(e.x), R.t = e.y, R.rng.s_hi = __float_as_uint(e.w);
    const uint32_t meta = __float_as_uint(e.z);
    R.state = (int)(meta & 7u), R.ncode = (int)((meta >> 3) & 7u), R.bounces = meta >> 6;
    R.rng.have = 0;  // a block never outlives its batch
  }
  CVR_DEV void store_static(unsigned i, const PathRegs<PhiloxCB>& R) {
    q[i] = make_float4(R.o.x, R.o.y, R.o.z, R.dist);
    q[W + i] = make_float4(R.d.x, R.d.y, R.d.z, __uint_as_float(R.rng.s_lo));
    q[2 * W + i] = make_float4(R.thr_x, R.thr_y, R.thr_z, __uint_as_float(R.out_idx));
  }
  CVR_DEV void store_dynamic(unsigned i, float t, uint32_t meta, const PhiloxCB& g) {
    q[3 * W + i] = make_float4(__uint_as_float(g.ctr), t, __uint_as_float(meta), __uint_as_float(g.s_hi));
  }
  CVR_DEV void store(unsigned i, const PathRegs<PhiloxCB>& R) {
    store_static(i, R);
    store_dynamic(i, R.t, (uint32_t)R.state | ((uint32_t)R.ncode << 3) | (R.bounces << 6), R.rng);
  }
  CVR_DEV void load_track(unsigned i, PathRegs<PhiloxCB>& R) {
    const float4 a = q[i], b = q[W + i], e = q[3 * W + i];
    R.o = v3(a.x, a.y, a.z), R.dist = a.w;
    R.d = v3(b.x, b.y, b.z), R.rng.s_lo = __float_as_uint(b.w);
    unpack_dynamic(e, R);
  }
  CVR_DEV void load(unsigned i, PathRegs<PhiloxCB>& R) {
    const float4 a = q[i], b = q[W + i], c = q[2 * W + i], e = q[3 * W + i];
    R.o = v3(a.x, a.y, a.z), R.dist = a.w;
    R.d = v3(b.x, b.y, b.z), R.rng.s_lo = __float_as_uint(b.w);
    R.thr_x = c.x, R.thr_y = c.y, R.thr_z = c.z, R.out_idx = __float_as_uint(c.w);
    R.path_lo = R.rng.s_lo - path_lo_base;
    unpack_dynamic(e, R);
  }
};
template <int RNGM>
struct WarpRngSel {
  typedef Xorwow type;
};
template <>
struct WarpRngSel<RNG_PHILOX> {
  typedef PhiloxCB type;
};
inline size_t warp_slot_bytes(int rng_mode) { return rng_mode == RNG_PHILOX ? 64 : sizeof(PathSlot); }

CVR_DEV int sort_key(int state) {
  // TRACK 0, SCATTER 1, BOUNDARY / BOUNDARY_P 2, IDLE 3, everything else (DONE) 4: a packed
  // 3-bit table indexed by the state (IDLE 0, ISECT 1, TRACK 2, SCATTER 3, BOUNDARY 4, DONE 5, BOUNDARY_P 6)
  constexpr uint32_t kTable = 3u | (4u << 3) | (0u << 6) | (1u << 9) | (2u << 12) | (4u << 15) | (2u << 18) | (4u << 21);
  return (int)((kTable >> (3u * (uint32_t)state)) & 7u);
}

template <int RNGM, int LAYOUT, bool COUNT>
__global__ void __launch_bounds__(CVR_BLOCK, CVR_MIN_BLOCKS)
    k_volpt_sorted(const __grid_constant__ KernelParams P) {
  typedef Xorwow Rng;
  constexpr int NW = CVR_BLOCK / 32;
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned warp = threadIdx.x >> 5;

  __shared__ PathSlot s_slot[CVR_BLOCK];
  __shared__ uint16_t s_list[CVR_BLOCK];
  __shared__ uint8_t s_state[CVR_BLOCK];
  __shared__ uint32_t s_cnt[5 * NW];  // per (key, warp) counts -> exclusive offsets
  __shared__ int s_exhausted;

  LaneCounters C;
  const unsigned long long per_tile = P.path_end - P.path_begin;
  const unsigned long long total = per_tile * P.n_launch_tiles;
  const TrackInv& I = P.inv;

  {  // all slots start idle
    PathRegs<Rng> R;
    R.o = v3(0, 0, 0), R.d = v3(0, 0, 1);
    R.thr_x = R.thr_y = R.thr_z = 1.f;
    R.t = R.dist = 0.f;
    R.out_idx = R.path_lo = R.bounces = 0;
    R.ncode = 0;
    R.state = S_IDLE;
    R.rng.init((int32_t)(P.seed + threadIdx.x + blockDim.x * blockIdx.x));  // per-slot stream (thread-rng mode)
    slot_store(s_slot[threadIdx.x], R);
    s_state[threadIdx.x] = S_IDLE;
    if (threadIdx.x == 0) s_exhausted = 0;
  }
  __syncthreads();

  for (;;) {
    // ---------------------------------------------------------------- sort slots by state
    const int my_key = sort_key(s_state[threadIdx.x]);
    uint32_t my_rank = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      unsigned m = __ballot_sync(FULL, my_key == k);
      if (my_key == k) my_rank = __popc(m & ((1u << lane) - 1u));
      if (lane == 0) s_cnt[k * NW + warp] = __popc(m);
    }
    __syncthreads();
    if (warp == 0) {
      // exclusive scan over the 5*NW counters in key-major order
      uint32_t a = lane < 5 * NW ? s_cnt[lane] : 0u;
      uint32_t b = (lane + 32) < 5 * NW ? s_cnt[lane + 32] : 0u;
      uint32_t ia = a, ib = b;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        uint32_t ta = __shfl_up_sync(FULL, ia, s), tb = __shfl_up_sync(FULL, ib, s);
        if ((int)lane >= s) ia += ta, ib += tb;
      }
      uint32_t tot_a = __shfl_sync(FULL, ia, 31);
      if (lane < 5 * NW) s_cnt[lane] = ia - a;
      if ((lane + 32) < 5 * NW) s_cnt[lane + 32] = tot_a + ib - b;
    }
    __syncthreads();
    s_list[s_cnt[my_key * NW + warp] + my_rank] = (uint16_t)threadIdx.x;
    // first index of the DONE group = number of live slots
    const uint32_t n_live = s_cnt[4 * NW];
    const uint32_t n_track = s_cnt[1 * NW];  // first index of the SCATTER group
    __syncthreads();
    if (n_live == 0) break;

    // ---------------------------------------------------------------- one round on my slot
    const unsigned slot = s_list[threadIdx.x];
    PathRegs<Rng> R;
    slot_load(s_slot[slot], R);
    bool exhausted = s_exhausted != 0;

    if (__any_sync(FULL, R.state != S_DONE)) {
      unsigned idle = __ballot_sync(FULL, R.state == S_IDLE);
      if (idle) {
        bool was = exhausted;
        warp_regenerate<RNGM, COUNT>(P, idle, lane, total, per_tile, exhausted, R, C);
        if (exhausted && !was && lane == 0) s_exhausted = 1;
      }
      if (R.state == S_SCATTER)
        do_scatter<LAYOUT, COUNT>(P, R, C);
      else if (R.state == S_BOUNDARY)
        do_boundary(P, R);
      if (R.state == S_ISECT) do_isect<COUNT>(P, R, C);
      // Woodcock steps; stop early when few lanes remain unless the CTA is draining
      const bool draining = n_track < 64u;
      for (int it = 0; it < P.track_steps; ++it) {
        unsigned trk = __ballot_sync(FULL, R.state == S_TRACK);
        if (trk == 0) break;
        if (it > 0 && !draining && __popc(trk) < P.track_min_lanes) break;
        if (R.state == S_TRACK) do_track_step<LAYOUT, COUNT>(P, I, R, C);
      }
      // states ISECT cannot persist across rounds
      slot_store(s_slot[slot], R);
      s_state[slot] = (uint8_t)R.state;
    }
    __syncthreads();
  }
  flush_counters<COUNT>(P, C, lane);
}

// =============================================================================
// Scheduler 3: decoupled state queues ("queued wavefront").  A CTA owns
// CVR_QSLOTS path slots in shared memory and one ring-buffer queue of slot ids per
// state (TRACK, SCATTER, BOUNDARY, IDLE).  There are NO block barriers in the steady
// state: every warp repeatedly pops up to 32 slot ids from the fullest queue, so its 32
// lanes hold paths in the SAME state, runs that state's event (+ the following
// intersect) and Woodcock steps while enough lanes stay busy, writes the paths back
// and pushes every slot onto the queue of its new state (warp-aggregated
// reserve -> write -> ordered commit).  Per-path operation and RNG order unchanged.
// =============================================================================
#ifndef CVR_QSLOTS
#define CVR_QSLOTS 512
#endif

struct QueueCtl {
  unsigned int head[4];  // pop cursor
  unsigned int tail[4];  // committed push cursor
  unsigned int resv[4];  // reserved push cursor
  unsigned int n_done;
  int exhausted;
};

CVR_DEV unsigned int ld_volatile_u32(const unsigned int* p) { return *(const volatile unsigned int*)p; }

// Write the lanes in `mask` back to their slots and enqueue every slot on the queue of
// its new state (warp-aggregated reserve -> write -> ordered commit); finished slots
// (S_DONE) are only counted.  Must be called by all 32 lanes.
template <unsigned N>
CVR_DEV void q_retire(uint16_t (*s_q)[N], QueueCtl& ctl, bool in_mask, unsigned slot, int state, unsigned lane) {
  const unsigned FULL = 0xffffffffu;
  const unsigned lane_lt = (1u << lane) - 1u;
  const int nk = in_mask ? sort_key(state) : 5;
  unsigned my_p = 0, commit_mask = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    unsigned m = __ballot_sync(FULL, nk == k);
    if (m) {
      int leader = __ffs(m) - 1;
      unsigned p = 0;
      if ((int)lane == leader) p = atomicAdd(&ctl.resv[k], (unsigned)__popc(m));
      p = __shfl_sync(FULL, p, leader);
      if (nk == k) s_q[k][(p + __popc(m & lane_lt)) & (N - 1)] = (uint16_t)slot;
      if ((int)lane == leader) my_p = p, commit_mask = m;
    }
  }
  unsigned done_m = __ballot_sync(FULL, nk == 4);
  // every lane's slot + queue stores must be visible before ANY leader publishes them:
  // fence my own stores, then order the warp (lanes do not run in lock step)
  __threadfence_block();
  __syncwarp();
  if (commit_mask) {
    // ordered commit: our entries become poppable after every earlier reservation
    const unsigned cnt = (unsigned)__popc(commit_mask);
    while (atomicCAS(&ctl.tail[nk], my_p, my_p + cnt) != my_p) {
    }
  }
  if (done_m && lane == 0) atomicAdd(&ctl.n_done, (unsigned)__popc(done_m));
  __syncwarp();
}

// Pop up to popc(want) slot ids from queue `key` into the lanes of `want` and load the
// paths.  The lanes READ their candidate entries first and only then lane 0 claims them
// with a CAS on the pop cursor: entries in [head, tail) are live and cannot be
// overwritten (a queue never holds more than N slots), so a successful CAS proves the
// ids read were the ones claimed.  (Claiming first and reading afterwards lets a fast
// warp wrap the ring over them.)  Returns the mask of lanes that received a path.
template <unsigned N>
CVR_DEV unsigned q_pop_into(const PathSlot* s_slot, uint16_t (*s_q)[N], QueueCtl& ctl, int key, unsigned want,
                            unsigned lane, unsigned& slot, PathRegs<Xorwow>& R) {
  const unsigned FULL = 0xffffffffu;
  const unsigned rank = __popc(want & ((1u << lane) - 1u));
  const bool wanted = (want >> lane) & 1u;
  for (int attempt = 0; attempt < 4; ++attempt) {
    unsigned h = 0, c = 0;
    if (lane == 0) {
      h = ld_volatile_u32(&ctl.head[key]);
      c = ld_volatile_u32(&ctl.tail[key]) - h;
    }
    h = __shfl_sync(FULL, h, 0);
    c = __shfl_sync(FULL, c, 0);
    unsigned n = min(c, (unsigned)__popc(want));
    if (n == 0) return 0u;
    __threadfence_block();
    const bool take = wanted && rank < n;
    unsigned cand = 0;
    if (take) cand = *(volatile uint16_t*)&s_q[key][(h + rank) & (N - 1)];
    int ok = 0;
    if (lane == 0) ok = atomicCAS(&ctl.head[key], h, h + n) == h;
    if (__shfl_sync(FULL, ok, 0)) {
      __threadfence_block();
      if (take) {
        slot = cand;
        if (key == 0)
          slot_load_track(s_slot[slot], R);
        else
          slot_load(s_slot[slot], R);
      }
      return __ballot_sync(FULL, take);
    }
  }
  return 0u;
}

template <int RNGM, int LAYOUT, bool COUNT, bool FAST, bool LOCAL = false>
__global__ void __launch_bounds__(CVR_BLOCK, CVR_MIN_BLOCKS)
    k_volpt_queued(const __grid_constant__ KernelParams P) {
  typedef Xorwow Rng;
  constexpr unsigned N = CVR_QSLOTS;
  static_assert((N & (N - 1)) == 0, "CVR_QSLOTS must be a power of two");
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;

  __shared__ PathSlot s_slot[N];
  __shared__ uint16_t s_q[4][N];
  __shared__ QueueCtl s_ctl;

  LaneCounters C;
  const unsigned long long per_tile = P.path_end - P.path_begin;
  const unsigned long long total = per_tile * P.n_launch_tiles;
  const TrackInv& I = P.inv;

  // all slots start idle, queued on the IDLE queue (key 3)
  for (unsigned i = threadIdx.x; i < N; i += blockDim.x) {
    PathRegs<Rng> R;
    R.o = v3(0, 0, 0), R.d = v3(0, 0, 1);
    R.thr_x = R.thr_y = R.thr_z = 1.f;
    R.t = R.dist = 0.f;
    R.out_idx = R.path_lo = R.bounces = 0;
    R.ncode = 0;
    R.state = S_IDLE;
    R.rng.init((int32_t)(P.seed + i + N * blockIdx.x));  // per-slot stream (thread-rng mode)
    slot_store(s_slot[i], R);
    s_q[3][i] = (uint16_t)i;
  }
  if (threadIdx.x == 0) {
    for (int k = 0; k < 4; ++k) s_ctl.head[k] = s_ctl.tail[k] = s_ctl.resv[k] = 0u;
    s_ctl.tail[3] = s_ctl.resv[3] = N;
    s_ctl.n_done = 0u;
    s_ctl.exhausted = 0;
  }
  __syncthreads();

  for (;;) {
    // ---------------------------------------------------------------- choose a queue
    // lanes 0..3 each inspect one queue; the fullest wins (ties -> lowest key).
    // (Measured alternatives, hetvol 1024^2: "full event batches first" -11 %,
    // in-place refill of finished tracking lanes -28 %: see DESIGN.md.)
    int key = -1;
    {
      unsigned c = 0;
      if (lane < 4) c = ld_volatile_u32(&s_ctl.tail[lane]) - ld_volatile_u32(&s_ctl.head[lane]);
      unsigned c0 = __shfl_sync(FULL, c, 0), c1 = __shfl_sync(FULL, c, 1), c2 = __shfl_sync(FULL, c, 2),
               c3 = __shfl_sync(FULL, c, 3);
      unsigned best = c0;
      key = c0 ? 0 : -1;
      if (c1 > best) best = c1, key = 1;
      if (c2 > best) best = c2, key = 2;
      if (c3 > best) best = c3, key = 3;
    }
    if (key < 0) {
      unsigned nd = lane == 0 ? ld_volatile_u32(&s_ctl.n_done) : 0u;
      if (__shfl_sync(FULL, nd, 0) >= N) break;
      __nanosleep(100);
      continue;
    }
    unsigned slot = 0;
    PathRegs<Rng> R;
    R.state = S_DONE;
    unsigned got = q_pop_into<N>(s_slot, s_q, s_ctl, key, FULL, lane, slot, R);
    if (!got) continue;  // somebody else emptied it first: start over
    bool have = (got >> lane) & 1u;

    // ---------------------------------------------------------------- event of this batch
    if (key == 3) {
      unsigned idle = __ballot_sync(FULL, have && R.state == S_IDLE);
      bool exhausted = ld_volatile_u32((const unsigned int*)&s_ctl.exhausted) != 0u;
      bool was = exhausted;
      if (idle) warp_regenerate<RNGM, COUNT>(P, idle, lane, total, per_tile, exhausted, R, C);
      if (exhausted && !was && lane == 0) s_ctl.exhausted = 1;
    } else if (key == 1) {
      if (have) do_scatter<LAYOUT, COUNT, FAST>(P, R, C);
    } else if (key == 2) {
      if (have) do_boundary<FAST>(P, R);
    }
    if (have && R.state == S_ISECT) do_isect<COUNT, FAST>(P, R, C);
    // everything the tracking loop does not touch goes back to the slot now
    if (key != 0 && have) slot_store_static(s_slot[slot], R);
    const uint32_t meta_hi = ((uint32_t)R.ncode << 3) | (R.bounces << 6);

    // ---------------------------------------------------------------- Woodcock steps
    if (FAST && LAYOUT != LAYOUT_LINEAR && !LOCAL) {
      const GridRay G = grid_ray(I, R.o, R.d);
      for (int it = 0; it < P.track_steps; it += P.pair ? 2 : 1) {
        unsigned trk = __ballot_sync(FULL, R.state == S_TRACK);  // lanes without a path are S_DONE
        if (trk == 0) break;
        if (it > 0 && __popc(trk) < P.track_min_lanes) {
          unsigned waiting = 0;
          if (lane == 0) waiting = ld_volatile_u32(&s_ctl.tail[0]) - ld_volatile_u32(&s_ctl.head[0]);
          if (__shfl_sync(FULL, waiting, 0) != 0u) break;
        }
        if (R.state == S_TRACK) {
          if (P.pair)
            track_pair_fast<LAYOUT, COUNT>(P, I, G, R, C);
          else
            track_step_fast<LAYOUT, COUNT>(P, I, G, R, C);
        }
      }
    } else {
      BrickWalk walk;  // local-majorant walk of this lane (located from scratch on entry)
      brick_walk_reset(walk);
      const GridRay G = grid_ray(I, R.o, R.d);
      for (int it = 0; it < P.track_steps; ++it) {
        unsigned trk = __ballot_sync(FULL, have && R.state == S_TRACK);
        if (trk == 0) break;
        if (it > 0 && __popc(trk) < P.track_min_lanes) {
          // few lanes left: requeue them so they merge into a full batch -- unless
          // nobody else is queued for tracking
          unsigned waiting = 0;
          if (lane == 0) waiting = ld_volatile_u32(&s_ctl.tail[0]) - ld_volatile_u32(&s_ctl.head[0]);
          if (__shfl_sync(FULL, waiting, 0) != 0u) break;
        }
        if (have && R.state == S_TRACK) {
          if (LOCAL)
            do_track_step_local<LAYOUT, COUNT>(P, I, G, R, C, walk);
          else
            do_track_step<LAYOUT, COUNT, FAST>(P, I, R, C);
        }
      }
    }

    // ---------------------------------------------------------------- write back + push
    if (have) slot_store_dynamic(s_slot[slot], R.t, meta_hi | (uint32_t)R.state, R.rng);
    q_retire<N>(s_q, s_ctl, have, slot, R.state, lane);
  }
  flush_counters<COUNT>(P, C, lane);
}

// =============================================================================
// Scheduler 4: warp-private wavefront ("warp").  ncu on the queued kernel (hetvol,
// profiles/r1_scheduler_evolution.md) shows 39 % of all issued instructions in queue
// management: four shared ring buffers per CTA with reserve / write / fence / ordered-CAS
// commit / CAS pop.  Nothing of that is needed if a warp never shares paths: here every
// warp owns CVR_WSLOTS path slots in shared memory and keeps the sort key of slot
// (lane + 32 j) in register st[j].  One round =
//   * ONE warp reduction (REDUX) of a packed byte-counter word gives the number of paths
//     per state; the fullest state wins;
//   * ballots + prefix popcounts rank the slots of that state; the first 32 write their id
//     to a 32-byte list (same warp reads it back after a __syncwarp) -> lane i runs slot
//     list[i], all lanes in the SAME state;
//   * event (+ intersect) + Woodcock steps exactly as in the queued kernel;
//   * the new key goes to a byte array the owner lanes read back.
// No atomics (except the global path counter at regeneration), no fences, no CAS, no
// block barriers; ~60 instructions of scheduling per batch instead of ~370.
// Each path's own operation and RNG order is unchanged, so results are identical to the
// other schedulers (event counters bit-identical, images up to fp32 atomic order).
// =============================================================================
// CTA shape of the warp scheduler: warps never interact, so the CTA size only sets the
// register budget granularity.  128 threads x 7 CTAs/SM = 28 warps at 72 registers (no spills)
// against 256 x 3 = 24 warps at 80: hetvol +2.4 %, manix +6.6 %, bucky +6.0 %, fbm 512^3 +0.6 %;
// 128 x 8 (64 registers, spills) -12 %.
#ifndef CVR_WBLOCK
#define CVR_WBLOCK 128
#endif
#ifndef CVR_WMIN_BLOCKS
#define CVR_WMIN_BLOCKS 7
#endif
enum : uint32_t { K_TRACK = 0, K_SCATTER = 1, K_BOUNDARY = 2, K_IDLE = 3, K_DONE = 4, K_BUSY = 5 };

// dynamic shared memory of k_volpt_warp for a CTA of `block` threads with W slots per warp
inline size_t warp_sched_smem_bytes(int block, int W, size_t skip_table_bytes = 0, size_t slot_bytes = sizeof(PathSlot)) {
  return (size_t)(block / 32) * (W * (slot_bytes + 1) + 32) + skip_table_bytes;
}

// (Measured and rejected, twice: in-place refill of finished tracking lanes from waiting
// paths of the same warp instead of ending the batch -- hetvol 930 -> 842, bucky 4798 -> 4190
// Msamples/s at the best refill period; profiles/r1_scheduler_evolution.md.)
// W = path slots per warp: more slots -> fuller batches but less L1 next to the slots
// (64: 145 KB of slots per SM at 28 warps, 96: 218 KB).  Measured on B200 (1024^2 x 16 spp,
// Msamples/s, W = 64 / 96): hetvol 889 / 961, bucky 5031 / 5276 (volumes resident in L2:
// fill wins), manix 2355 / 2102, fbm 512^3 1059 / 1031 (volumes beyond L2: L1 wins).
// SKIP = fetch-skip table (SkipTab above) in shared memory behind the slots.  It has to be ONE
// table per SM to leave room for the slots, so these instantiations run as one CTA of
// CVR_WSKIP_BLOCK = 896 threads per SM (the same 28 warps and 72 registers as 7 x 128; warps
// never interact, the CTA only shares the table and one barrier after loading it).
#ifndef CVR_WSKIP_BLOCK
#define CVR_WSKIP_BLOCK 896
#endif
// LOG = the per-path event log of cvr_trace_paths_logged (parity hook; its own instantiations, so
// that the product kernels carry no trace of it: with a run-time test alone ptxas spilled 16-38
// bytes in the event code of the 72-register kernels)
template <int RNGM, int LAYOUT, bool COUNT, bool FAST, bool LOCAL = false, int W = 64, bool SKIP = false, bool LOG = false>
__global__ void __launch_bounds__(SKIP ? CVR_WSKIP_BLOCK : CVR_WBLOCK, SKIP ? 1 : CVR_WMIN_BLOCKS)
    k_volpt_warp(const __grid_constant__ KernelParams P) {
  static_assert(!SKIP || (FAST && !LOCAL && LAYOUT != LAYOUT_LINEAR), "the skip table belongs to the fused global-majorant loop");
  typedef typename WarpRngSel<RNGM>::type Rng;
  typedef WarpSlots<Rng, W> Slots;
  constexpr bool CB = RNGM == RNG_PHILOX;  // counter-based stream: 64-byte slots, no roll-back, no parked uniform
  static_assert(!CB || !LOG, "the event log records XORWOW draw counters");
  constexpr size_t SB = Slots::kSlotBytes;
  constexpr int K = W / 32;
  static_assert(W % 32 == 0 && K >= 1 && K <= 7, "slots per warp must be 32..224 in steps of 32");
  const unsigned FULL = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned lane_lt = (1u << lane) - 1u;

  extern __shared__ __align__(16) unsigned char s_raw[];
  Slots slots(s_raw + (size_t)warp * W * SB, (uint32_t)(P.seed + (uint32_t)P.path_begin));
  uint8_t* const keys = s_raw + (size_t)nw * W * SB + warp * W;
  uint8_t* const list = s_raw + (size_t)nw * W * (SB + 1) + warp * 32;
  // SKIP kernels always run with CVR_WSKIP_BLOCK threads: the table offset is a compile-time constant
  constexpr size_t kTabOffset = (size_t)(CVR_WSKIP_BLOCK / 32) * (W * (SB + 1) + 32);
  SkipTab ST;
  ST.tab = s_raw + kTabOffset;
  if (SKIP) {  // stage the table once per CTA
    uint8_t* const tab = s_raw + kTabOffset;
    for (uint32_t i = threadIdx.x * 4u; i < P.skip_n; i += blockDim.x * 4u)  // skip_n is padded to 4 bytes
      *reinterpret_cast<uint32_t*>(tab + i) = __ldg(reinterpret_cast<const uint32_t*>(P.skip_tab + i));
    __syncthreads();
  }

  LaneCounters C;
  const unsigned long long per_tile = P.path_end - P.path_begin;
  const unsigned long long total = per_tile * P.n_launch_tiles;
  const TrackInv& I = P.inv;

  uint32_t st[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {  // all slots start idle
    PathRegs<Rng> R;
    R.o = v3(0, 0, 0), R.d = v3(0, 0, 1);
    R.thr_x = R.thr_y = R.thr_z = 1.f;
    R.t = R.dist = 0.f;
    R.out_idx = R.path_lo = R.bounces = 0;
    R.ncode = 0;
    R.state = S_IDLE;
    // per-slot stream (thread-rng mode)
    if constexpr (CB)
      R.rng.init(0ull);
    else
      R.rng.init((int32_t)(P.seed + (blockIdx.x * nw + warp) * W + lane + 32 * j));
    slots.store(lane + 32 * j, R);
    st[j] = K_IDLE;
  }
  __syncwarp();
  bool exhausted = false;

  for (;;) {
    // ---------------------------------------------------------------- census + choice
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) packed += st[j] < 4u ? (1u << (8u * st[j])) : 0u;
    const uint32_t tot = __reduce_add_sync(FULL, packed);  // 4 byte counters, each <= 32 K
    if (tot == 0u) break;                                  // every slot is DONE
    const uint32_t c0 = tot & 255u, c1 = (tot >> 8) & 255u, c2 = (tot >> 16) & 255u, c3 = tot >> 24;
    uint32_t key = 0, best = c0;
    if (P.policy == 1 && c0 < 32u) best = 0;  // not a full tracking batch: run events first
    if (c1 > best) best = c1, key = 1;
    if (c2 > best) best = c2, key = 2;
    if (c3 > best) best = c3, key = 3;

    // ---------------------------------------------------------------- pick <= 32 slots of that state
    // (Measured and rejected: ranking the slots within their bank-residue class -- slot s has its
    // 16-byte fields in bank group (5 s + k) mod 8, so a quarter-warp holding 8 different residues
    // is conflict-free -- removes 67 % of the shared-memory bank conflicts and halves the
    // long-scoreboard stall, but a class with fewer than 4 candidates leaves holes in the batch:
    // +16 % instructions, hetvol 1098 -> 1012, bucky 5620 -> 5131 Msamples/s; restricted to states
    // with >= 40 candidates it is neutral; with a second pass that fills the holes with the
    // left-over candidates through a temporary list only 36 % of the conflicts go and the extra
    // ranking costs more than they did: hetvol 1210 -> 1174, bucky 5979 -> 5570.)
    unsigned base = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
      const bool mine = st[j] == key;
      const unsigned m = __ballot_sync(FULL, mine);
      const unsigned r = base + __popc(m & lane_lt);
      if (mine && r < 32u) {
        list[r] = (uint8_t)(lane + 32 * j);
        keys[lane + 32 * j] = (uint8_t)K_BUSY;
        st[j] = K_BUSY;
      }
      base += __popc(m);
    }
    __syncwarp();
    const unsigned n = min(base, 32u);
    const bool have = lane < n;
    unsigned slot = 0;
    PathRegs<Rng> R;
    R.state = S_DONE, R.ncode = 0, R.bounces = 0;
    if (have) {
      slot = list[lane];
      if (key == 0)
        slots.load_track(slot, R);
      else
        slots.load(slot, R);
    }

    // ---------------------------------------------------------------- event of this batch
    if (key == 3) {
      unsigned idle = __ballot_sync(FULL, have && R.state == S_IDLE);
      if (idle) warp_regenerate<RNGM, COUNT>(P, idle, lane, total, per_tile, exhausted, R, C);
    } else if (key == 1) {
      if (have) do_scatter<LAYOUT, COUNT, FAST, LOG>(P, R, C);
    } else if (key == 2) {
      if (have) do_boundary<FAST, LOG, SKIP>(P, R);  // SKIP: small-argument sin / cos (cvr_device.cuh)
    }
    if (have && R.state == S_ISECT) do_isect<COUNT, FAST, LOG>(P, R, C);
    // everything the tracking loop does not touch goes back to the slot now
    if (key != 0 && have) slots.store_static(slot, R);
    const uint32_t meta_hi = ((uint32_t)R.ncode << 3) | (R.bounces << 6);

    // ---------------------------------------------------------------- Woodcock steps
    // tracking paths of this warp that are NOT in this batch: worth leaving the loop early
    // for (they merge with the stragglers into a fuller batch)
    // ... and also when enough of the warp's other slots wait for ANY batch (P.exit_others, default
    // 16): with short segments (fBm, albedo 0.99: 4 lookups per segment) a few stragglers would
    // otherwise step alone for the rest of the batch while most paths sit in scatter state.
    // fbm 512^3 1497 -> 1580 Msamples/s, bucky / hetvol / manix / sparse within +-0.5 %.
    const uint32_t others = (c0 + c1 + c2 + c3) - n;
    const bool others_track = (key == 0 ? c0 - n : c0) != 0u || (P.exit_others > 0 && others >= (uint32_t)P.exit_others);
    if (FAST && LAYOUT != LAYOUT_LINEAR && !LOCAL) {
      const GridRay G = grid_ray(I, R.o, R.d);
      // few lanes left and other tracking paths of this warp waiting in slots: stop, they merge
      const int steps = P.track_steps, min_lanes = others_track ? P.track_min_lanes : 0;
      if constexpr (CB) {
        // counter-based stream: always the pair step, one Philox block per pair (track_pair_cb)
        int need = 1;
        for (int it = 0; it < steps; it += 2) {
          const unsigned trk = __ballot_sync(FULL, R.state == S_TRACK);
          if (__popc(trk) < need) break;
          need = max(min_lanes, 1);
          if (R.state == S_TRACK) track_pair_cb<LAYOUT, COUNT, SKIP>(P, I, G, R, C, ST);
        }
#ifndef CVR_SKIP_FORCE_PAIR
#define CVR_SKIP_FORCE_PAIR 0
#endif
      } else if ((CVR_SKIP_FORCE_PAIR && SKIP) || P.pair) {
        // one test per iteration: stop below `need` tracking lanes; the first iteration runs with any.
        // (Two pairs per vote -- the census is 15 of the ~165 instructions of an iteration and runs
        // with all 32 lanes -- was measured and rejected: the coarser exit costs more than the votes,
        // bucky 5979 -> 5461, hetvol 1211 -> 1177, manix 2907 -> 2658 Msamples/s.)
        int need = 1;
        for (int it = 0; it < steps; it += 2) {
          const unsigned trk = __ballot_sync(FULL, R.state == S_TRACK);  // lanes without a path are S_DONE
          if (__popc(trk) < need) break;
          need = max(min_lanes, 1);
          if (R.state == S_TRACK) track_pair_fast<LAYOUT, COUNT, SKIP>(P, I, G, R, C, ST);
        }
      } else {
        for (int it = 0; it < steps; ++it) {
          const unsigned trk = __ballot_sync(FULL, R.state == S_TRACK);
          if (trk == 0 || (it > 0 && __popc(trk) < min_lanes)) break;
          if (R.state == S_TRACK) track_step_fast<LAYOUT, COUNT, SKIP>(P, I, G, R, C, ST);
        }
      }
    } else {
      BrickWalk walk;  // local-majorant walk of this lane (located from scratch on entry)
      brick_walk_reset(walk);
      const GridRay G = grid_ray(I, R.o, R.d);
      for (int it = 0; it < P.track_steps; ++it) {
        unsigned trk = __ballot_sync(FULL, R.state == S_TRACK);
        if (trk == 0) break;
        if (it > 0 && others_track && __popc(trk) < P.track_min_lanes) break;
        if (R.state == S_TRACK) {
          if (LOCAL)
            do_track_step_local<LAYOUT, COUNT>(P, I, G, R, C, walk);
          else
            do_track_step<LAYOUT, COUNT, FAST>(P, I, R, C);
        }
      }
    }

    // ---------------------------------------------------------------- write back + new keys
    if (have) {
      slots.store_dynamic(slot, R.t, meta_hi | (uint32_t)R.state, R.rng);
      keys[slot] = (uint8_t)sort_key(R.state);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (st[j] == K_BUSY) st[j] = keys[lane + 32 * j];
  }
  flush_counters<COUNT, SKIP>(P, C, lane);
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
    // corner order of cvr_device.cuh: (x1,y1,z1) (x1,y1,z2) (x2,y1,z1) (x2,y1,z2) | same for y2
    a.x = D[X1 + sx * Y1 + sxy * Z1], a.y = D[X1 + sx * Y1 + sxy * Z2];
    a.z = D[X2 + sx * Y1 + sxy * Z1], a.w = D[X2 + sx * Y1 + sxy * Z2];
    b.x = D[X1 + sx * Y2 + sxy * Z1], b.y = D[X1 + sx * Y2 + sxy * Z2];
    b.z = D[X2 + sx * Y2 + sxy * Z1], b.w = D[X2 + sx * Y2 + sxy * Z2];
    cells[2 * i] = a;
    cells[2 * i + 1] = b;
  }
}

__global__ void k_build_albedo_cells(const float4* __restrict__ A, int nx, int ny, int nz,
                                     float* __restrict__ cells) {
  size_t ncell = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
  // 8 threads per cell, one corner each: (r, g) to floats 2 i, 2 i + 1 and b to 16 + its
  // density-order index (AlbedoCell, cvr_device.cuh); the 96 bytes of a cell are written by one
  // group of 8 consecutive lanes
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
    const float4 a = A[X + (size_t)nx * (Y + (size_t)ny * Z)];
    float* cell = cells + c * CVR_ACELL_FLOATS;
    *reinterpret_cast<float2*>(cell + 2 * corner) = make_float2(a.x, a.y);
    cell[16 + (((corner >> 2) & 1) | ((corner & 1) << 1) | (((corner >> 1) & 1) << 2))] = a.z;
  }
}

__global__ void k_build_majorant(const float4* __restrict__ cells, int nx, int ny, int nz, uint32_t mx,
                                 uint32_t my, uint32_t mz, float* __restrict__ maj) {
  // one warp per brick, lanes stride over its cells
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  const uint32_t n_bricks = mx * my * mz;
  if (warp >= n_bricks) return;
  const uint32_t bx = warp % mx, by = (warp / mx) % my, bz = warp / (mx * my);
  float m = 0.f;
  for (uint32_t i = lane; i < CVR_BRICK * CVR_BRICK * CVR_BRICK; i += 32) {
    uint32_t kx = bx * CVR_BRICK + (i % CVR_BRICK), ky = by * CVR_BRICK + ((i / CVR_BRICK) % CVR_BRICK),
             kz = bz * CVR_BRICK + i / (CVR_BRICK * CVR_BRICK);
    if (kx > (uint32_t)nx || ky > (uint32_t)ny || kz > (uint32_t)nz) continue;
    size_t cell = kx + (size_t)(nx + 1) * (ky + (size_t)(ny + 1) * kz);
    float4 a = cells[2 * cell], b = cells[2 * cell + 1];
    m = fmaxf(m, fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w))));
  }
  for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
  if (lane == 0) maj[warp] = m;
}

// second majorant level: max over 8^3 bricks
__global__ void k_build_majorant2(const float* __restrict__ maj, uint32_t mx, uint32_t my, uint32_t mz,
                                  float* __restrict__ maj2, uint32_t m2x, uint32_t m2y, uint32_t m2z) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m2x * m2y * m2z) return;
  const uint32_t cx = i % m2x, cy = (i / m2x) % m2y, cz = i / (m2x * m2y);
  float m = 0.f;
  for (uint32_t z = 8 * cz; z < min(8 * cz + 8, mz); ++z)
    for (uint32_t y = 8 * cy; y < min(8 * cy + 8, my); ++y)
      for (uint32_t x = 8 * cx; x < min(8 * cx + 8, mx); ++x) m = fmaxf(m, maj[x + (size_t)mx * (y + (size_t)my * z)]);
  maj2[i] = m;
}

// ---------------------------------------------------------------- resolve (A16)
// ImageBufferTransfer.cu:6-18 with UtilityFunctors::Scale (Utilities.h:6-15): every
// float of the tile divided by `scale`, written at the tile origin of the image.
// fetch-skip table (SkipTab): byte per brick of (8 << e)^3 cells from the 8^3-brick majorants
__global__ void k_build_skip_table(const float* __restrict__ maj, uint32_t mx, uint32_t my, uint32_t mz, uint32_t e,
                                   uint32_t bx, uint32_t by, uint32_t bz, float sig_ratio, uint8_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bx * by * bz) return;
  const uint32_t X = i % bx, Y = (i / bx) % by, Z = i / (bx * by);
  float m = 0.f;
  for (uint32_t z = Z << e; z < min((Z + 1u) << e, mz); ++z)
    for (uint32_t y = Y << e; y < min((Y + 1u) << e, my); ++y)
      for (uint32_t x = X << e; x < min((X + 1u) << e, mx); ++x) m = fmaxf(m, maj[x + (size_t)mx * (y + (size_t)my * z)]);
  // r = the largest value `density * sig_ratio` can take in the brick (same fp32 multiply as the loop)
  const float r = m * sig_ratio;
  float q = ceilf(r * 256.0f * 1.00001f);
  q = !(q >= 1.0f) ? 1.0f : (q > 256.0f ? 256.0f : q);  // NaN / negative -> never more permissive than r = 1/256
  if (!(r <= 1.0f)) q = 256.0f;                           // NaN or majorant above max_density: never skip
  out[i] = (uint8_t)((int)q - 1);
}

// After the framebuffers of several devices were summed: the alpha channel is not a sum.  A path that escapes STORES
// w = 1 (Utilities.cuh:15-22, Q12) and the resolve divides by the iteration count, so a pixel's alpha is 1 / iterations
// if any of its paths escaped and 0 otherwise; sample-sharded ranks each contribute that value, and the sum counts the
// ranks.  min(sum, 1 / iterations) restores exactly what one device produces.
__global__ void k_clamp_alpha(float4* __restrict__ image, size_t n_px, float limit) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_px; i += (size_t)gridDim.x * blockDim.x)
    image[i].w = fminf(image[i].w, limit);
}

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

// The fused-tiles form of the resolve: ONE launch over every tile of the table (blockIdx.y = tile number), tiles this
// rank does not own (up to two interleave phases, cvr_shard) are skipped.  100 launches of k_resolve_tile per C3 image
// were 1.0-1.5 ms of host time per render -- serialised on the driver lock when 8 host threads render for a device group.
struct ResolveOwner {
  uint32_t n, first[2], stride[2], limit[2];
};
// ADD = the resolve fused with the cross-device sum: instead of storing into its own image, every device of a group ADDS
// its resolved share straight into the FIRST device's image through NVLink peer memory -- one 16-byte
// `red.relaxed.sys.global.add.v4.f32` per pixel (SASS REDG.E.ADD.F32x4...SYS), fire and forget.  No per-device
// full-resolution image, no zero fill of it, no collective call: the sum happens in the first device's L2.
template <bool ADD>
__global__ void k_resolve_tiles(const float4* __restrict__ accum, uint32_t tile_w, uint32_t tile_h, uint32_t full_w,
                                const uint2* __restrict__ origins, float4* __restrict__ image, float scale, ResolveOwner own) {
  const uint32_t k = blockIdx.y;
  bool mine = false;
  for (uint32_t p = 0; p < own.n; ++p)
    mine |= k >= own.first[p] && k < own.limit[p] && (k - own.first[p]) % own.stride[p] == 0;
  if (!mine) return;
  const uint2 o = origins[k];
  const uint32_t n = tile_w * tile_h;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t x = i % tile_w, y = i / tile_w;
    const size_t at = (size_t)(y + o.y) * full_w + (x + o.x);
    float4 v = accum[at];
    v.x = v.x / scale, v.y = v.y / scale, v.z = v.z / scale, v.w = v.w / scale;
    if (ADD)
      asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(image + at), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                   : "memory");
    else
      image[at] = v;
  }
}

// ---------------------------------------------------------------- display resolve (progressive / interactive path)
// DeviceTiledImageBufferTansferDelegate::transfer (ImageBufferTransfer.cu:20-59, 128-157) with
// ColorPixelTransform<Scale> (:80-100): the tile's accumulation buffer is ADDED into a full-resolution
// float4 transfer buffer at the tile origin (negative and NaN contributions count as 0, alpha is left
// alone), and the running sum goes out as 8-bit display pixels: c = clamp(pow(sum / scale, 1 / 2.2) *
// 255, 0, 255) truncated, alpha 255.  One thread per pixel, rows coalesced (16 B in, 16 B read-modify-
// write, 4 B out).
CVR_DEV unsigned char display_channel(float v, float scale) {
  const float g = __powf(v / scale, 1.f / 2.2f);
  return (unsigned char)fminf(fmaxf(g * 255.f, 0.f), 255.f);
}
__global__ void k_accumulate_display(const float4* __restrict__ tile, uint32_t tile_w, uint32_t tile_h, float4* __restrict__ transfer,
                                     uchar4* __restrict__ display, uint32_t full_w, uint32_t full_h, uint32_t off_x, uint32_t off_y,
                                     float scale) {
  const uint32_t n = tile_w * tile_h;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t x = i % tile_w, y = i / tile_w;
    const uint32_t ox = x + off_x, oy = y + off_y;
    if (ox >= full_w || oy >= full_h) continue;  // ImageBufferTransfer.cu:36-38
    const float4 in = tile[i];
    const size_t o = (size_t)oy * full_w + ox;
    float4 t = transfer[o];
    t.x += in.x > 0 ? in.x : 0;
    t.y += in.y > 0 ? in.y : 0;
    t.z += in.z > 0 ? in.z : 0;
    transfer[o] = t;
    display[o] = make_uchar4(display_channel(t.x, scale), display_channel(t.y, scale), display_channel(t.z, scale), 255);
  }
}

// ---------------------------------------------------------------- gather roofline
// Random 32-byte-sector gather microbenchmark: every thread issues
// `per_thread` independent 256-bit loads at hashed cell indices of a `n_cells`-cell
// buffer, UNROLL in flight at a time; the sum defeats dead-code elimination.  This is
// the measured "gather roofline" denominator for a given footprint (L2-resident or HBM).
CVR_DEV void ldg256_nol1(const float* p, float (&v)[8]) {  // the same load, not allocated in the L1
  asm("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
      : "l"(p));
}
template <int UNROLL, bool NOL1 = false>
__global__ void __launch_bounds__(256) k_gather_bench(const float* __restrict__ cells, uint32_t n_cells,
                                                       int per_thread, float* sink) {
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  for (int i = 0; i < per_thread; i += UNROLL) {
    float v[UNROLL][8];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      x = x * 1664525u + 1013904223u;
      uint32_t h = x ^ (x >> 15);
      h *= 0x2c1b3c6du;
      h ^= h >> 12;
      uint32_t cell = (uint32_t)(((unsigned long long)h * n_cells) >> 32);
      if (NOL1)
        ldg256_nol1(cells + 8 * (size_t)cell, v[u]);
      else
        ldg256(cells + 8 * (size_t)cell, v[u]);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) acc += v[u][0] + v[u][7];
  }
  if (acc == 1.2345e-30f) *sink = acc;
}

// ---------------------------------------------------------------- debug kernels
// parity hook: sin_small / cos_small / tan_small against libdevice's sinf / cosf / tanf on EVERY float of magnitude <= limit
// (both signs) and on NaN; mismatch counts per function, the first mismatching bit pattern kept for the error message
__global__ void k_trig_check(uint32_t limit_bits, unsigned long long* mismatches, uint32_t* first_bad) {
  const uint32_t stride = gridDim.x * blockDim.x;
  unsigned long long bad[3] = {0, 0, 0};
  for (uint32_t m = blockIdx.x * blockDim.x + threadIdx.x; m <= limit_bits + 1u; m += stride) {
    for (int sgn = 0; sgn < 2; ++sgn) {
      const uint32_t bits = m > limit_bits ? (0x7FC00000u | (sgn << 31)) : (m | ((uint32_t)sgn << 31));  // last m: NaN
      const float x = __uint_as_float(bits);
      const float a[3] = {sin_small(x), cos_small(x), tan_small(x)};
      const float b[3] = {sinf(x), cosf(x), tanf(x)};
      for (int f = 0; f < 3; ++f) {
        const bool same = __float_as_uint(a[f]) == __float_as_uint(b[f]) || (a[f] != a[f] && b[f] != b[f]);
        if (!same) {
          ++bad[f];
          atomicMin(&first_bad[f], bits & 0x7FFFFFFFu);
        }
      }
    }
    if (m == 0xFFFFFFFFu) break;
  }
  for (int f = 0; f < 3; ++f)
    if (bad[f]) atomicAdd(&mismatches[f], bad[f]);
}

__global__ void k_rng_kat(const int32_t* seeds, int n_seeds, int n, uint32_t* words, float* uni) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_seeds) return;
  Xorwow a, b;
  a.init(seeds[i]);
  b.init(seeds[i]);
  for (int k = 0; k < n; ++k) {
    // draw, give the draw back, draw again: also pins Xorwow::undo against cuRAND's words
    uint32_t w0 = a.next_u32();
    a.undo();
    uint32_t w1 = a.next_u32();
    words[(size_t)i * n + k] = (w0 == w1) ? w1 : ~w1;
    uni[(size_t)i * n + k] = b.next();
  }
}

template <int LAYOUT>
__global__ void k_debug_lookup(MediumParams m, TrackInv inv, const float* p, int n, float* dens, float* alb) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  V3 c = v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
  dens[i] = density_at<LAYOUT>(m, inv, c);
  V3 a = albedo_lookup<LAYOUT>(m, c);
  alb[3 * i] = a.x, alb[3 * i + 1] = a.y, alb[3 * i + 2] = a.z;
}

}  // namespace cvr
