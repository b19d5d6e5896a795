/*
 * cvr_oracle.h -- CPU oracle for the CudaVolumeRenderer per-pixel path loop.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * estimator (reference = Fe0437/CudaVolumeRenderer, paths below are relative to
 * its implementation/src/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path
 * (cudavolumerenderer_b200/csrc) never links or calls it.
 *
 * Parity pin: the reference ships no tests, golden vectors or CPU renderer
 * (SURVEY.md section 8(c)), so this oracle is pinned against the reference ITSELF
 * run here, three ways (recipes in oracle/Makefile, outputs in oracle/_ref/):
 *  1. WHOLE PATHS, bit for bit: oracle/_ref/libcvr_ref_cpu.so is the reference's own
 *     NaiveVolPTsk_kernel::d_render and RegenerationVolPTsk_kernel::
 *     d_render_single_thread_regeneration -- with woodcockTracking, DeviceVolume::
 *     operator(), indexToCameraRay, AABB::intersect, GGX::sample, HG::sample,
 *     atomicVectorAdd -- compiled for the host by g++ from the headers where they lie
 *     (oracle/ref_cpu_harness.cpp supplies only what the CUDA platform provides:
 *     thread indices, atomicAdd, a host XORWOW behind the Rng interface, the point-sampled
 *     clamped texture fetch).  tests/test_oracle.py requires every per-path radiance and
 *     every single-stream regenerationSK image of this file to EQUAL that build's, live
 *     and against tests/golden/ref_cpu_paths.npz (generated from it, script committed).
 *  2. Per function: oracle/_ref/libcvr_ref_host.so host-compiles the reference's
 *     __host__ __device__ functions through nvcc (AABB::intersect, Frame,
 *     ImportanceSampleHG, GGX_sample, fresnelDielectric, GGX_G1, morton3D, utilhash);
 *     checked bit for bit live and against tests/golden/ref_host_golden.npz.
 *  3. On the GPU: the reference's own kernels compiled for sm_100a
 *     (oracle/_ref/libcvr_ref_gpu.so) and cuRAND's device XORWOW (tests/test_gpu_parity.py).
 */
#ifndef CVR_ORACLE_H_
#define CVR_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG: cuRAND XORWOW, seed-only init (Rng.h:22, curand_kernel.h) ---- */
typedef struct {
  uint32_t v[5];
  uint32_t d;
} cvro_rng;

void cvro_rng_init(cvro_rng* r, int32_t seed);
uint32_t cvro_rng_u32(cvro_rng* r);
float cvro_rng_float(cvro_rng* r); /* curand_uniform: (0,1] */

/* ---- scene / camera description (Medium.h:109-187, RenderKernelLauncher.cu:67-72) ---- */
typedef struct {
  const float* density; /* x-fastest, dnx*dny*dnz floats */
  int32_t dnx, dny, dnz;
  const float* albedo; /* x-fastest float4 (rgb + pad), anx*any*anz*4 floats */
  int32_t anx, any, anz;
  float box_min[3];
  float box_max[3];
  float scale;
  float max_density;
  float hg_g;
  float ggx_alpha[2];
  float ggx_eta; /* int_ior / ext_ior */
} cvro_scene;

typedef struct {
  float inv_view[12];         /* c_inv_view_mat: 3 rows x 4 */
  float raster_to_view[2];    /* c_raster_to_view */
  float resolution[2];        /* c_resolution: TILE size */
  float pixel_index_range[2]; /* c_pixel_index_range: FULL image size */
  uint32_t offset[2];         /* c_offset: tile origin */
} cvro_camera;

typedef struct {
  uint64_t paths;           /* paths started */
  uint64_t bounces;         /* loop iterations (= the thesis' "rays") */
  uint64_t density_lookups; /* A7 invocations */
  uint64_t albedo_lookups;  /* A8 invocations */
  uint64_t escaped;         /* paths that reached the environment */
} cvro_counters;

enum {
  CVRO_NAIVE = 0, /* NaiveVolPTsk_kernel.cuh:17-87 */
  CVRO_REGEN = 1  /* RegenerationVolPTsk_kernel.cuh:146-232 */
};

/* Trace ONE path with an already-initialised RNG.  Returns 1 and fills
 * radiance[3] when the path escaped to the environment, 0 when Russian roulette
 * ended it.  variant selects the scatter pull-back quirk (Q8).  max_bounces == 0
 * means unbounded, as in the reference. */
int cvro_trace_path(const cvro_scene* sc, const cvro_camera* cam, cvro_rng* rng,
                    uint32_t image_id, int variant, uint32_t max_bounces,
                    float radiance[3], cvro_counters* ctr);

/* ---- decision / event log of one path (observes only; see cvr_oracle.c) ----
 * events: one entry per loop iteration that ended in a scatter or boundary event, plus
 *   the final escape; ev_d = XORWOW draw counter `d` when the event starts (after the
 *   Woodcock loop), so two implementations with the same stream agree on (code, d).
 * decisions: every compare a rounding difference can flip, with the draw counter right
 *   after the draw involved and the two compared values. */
enum {
  CVRO_EV_SCATTER = 1, CVRO_EV_BOUNDARY = 2, CVRO_EV_ESCAPE = 3,
  CVRO_EVF_OK = 16,      /* boundary: the GGX sample succeeded */
  CVRO_EVF_WO_NEG = 32,  /* boundary: local wo.z < 0 (reflect / refract side) */
  CVRO_EVF_WI_NEG = 64,  /* boundary: local wi.z < 0 */
  CVRO_EVF_KILLED = 128, /* Russian roulette ended the path after this event */
  CVRO_EVF_ZERO = 256    /* the throughput is exactly zero after the event: GGX_G1 returns 0 when 1 - wo.z^2 <= 0
                          * (GGX.h:232-236), and wo is not renormalised after refract (GGX.h:45-50), so a
                          * near-axial refraction is weight 0 or ~1 depending on the last bit of wo.z */
};
enum {
  CVRO_DEC_EXIT = 1,     /* a = t, b = max_t          : continue while t <= max_t  */
  CVRO_DEC_ACCEPT = 2,   /* a = sigma_t/sigma_max, b = u' : null collision while a < b */
  CVRO_DEC_ROULETTE = 3, /* a = u, b = p_survive      : killed when a > b          */
  CVRO_DEC_FRESNEL = 4   /* a = u, b = F              : reflect when a <= b        */
};
typedef struct {
  uint32_t d0; /* draw counter right after Rng(seed) */
  uint32_t n_events, cap_events;
  uint32_t* ev_code;
  uint32_t* ev_d;
  uint32_t n_dec, cap_dec;
  uint32_t* dec_kind;
  uint32_t* dec_d;
  float* dec_a;
  float* dec_b;
} cvro_trace;

/* cvro_trace_path with Rng(rng_seed) and the log filled (counts may exceed the caps:
 * entries beyond are dropped).  Returns 1 when the path escaped. */
int cvro_trace_path_logged(const cvro_scene* sc, const cvro_camera* cam, int32_t rng_seed,
                           uint32_t image_id, int variant, uint32_t max_bounces,
                           float radiance[3], cvro_trace* tr);

/* Per-path radiances with Rng(seed + path id) and the given variant (pull-back quirk):
 * out_per_path[4*(p-first)] for p in [first, first+count). */
void cvro_trace_paths_seeded(const cvro_scene* sc, const cvro_camera* cam, uint64_t first,
                             uint64_t count, uint32_t seed, int variant, float* out_per_path,
                             int n_host_threads, cvro_counters* ctr);

/* naiveSK launch: path tid in [0, w*h*iterations), Rng(tid), image_id = tid % (w*h).
 * out = tile accumulation buffer (w*h float4), accumulated into, w set to 1 on
 * every pixel that received an escaped path (Utilities.cuh:15-22).
 * Sums per pixel are taken in ascending path order (the reference's atomics are
 * unordered). */
void cvro_render_naive(const cvro_scene* sc, const cvro_camera* cam,
                       uint32_t iterations, float* out, int n_host_threads,
                       cvro_counters* ctr);

/* Per-path variant of the above: out_per_path[4*p] for p in [first, first+count). */
void cvro_trace_paths_naive(const cvro_scene* sc, const cvro_camera* cam,
                            uint64_t first, uint64_t count, float* out_per_path,
                            int n_host_threads, cvro_counters* ctr);

/* regenerationSK launch.
 *  rng_mode 0 ("thread", the reference): n_persistent virtual threads, virtual
 *    thread v owns Rng(seed + v) for its whole life and claims paths v, v+T, ...
 *    (one valid realisation of the reference's unordered atomic queue, Q7); the
 *    Russian-roulette draw after an escape is consumed (Q8).
 *  rng_mode 1 ("path", the new build's reproducible default): path p owns
 *    Rng(seed + p). */
void cvro_render_regen(const cvro_scene* sc, const cvro_camera* cam,
                       uint32_t iterations, uint32_t seed, int rng_mode,
                       uint32_t n_persistent, float* out, int n_host_threads,
                       cvro_counters* ctr);

/* ---- pieces exposed for unit pinning against the reference's host-compiled code ---- */
int cvro_aabb_intersect(const float bmin[3], const float bmax[3],
                        const float o[3], const float d[3], float* dist,
                        float normal[3], int* inside);
void cvro_frame_from_z(const float n[3], float x[3], float y[3], float z[3]);
void cvro_hg_sample(const float dir[3], float g, float e1, float e2, float out[3]);
/* returns success; wo written exactly as the reference writes *wo (also on the
 * failing reflect/refract checks); draws come from u[3] in order, *n_used says
 * how many were consumed. */
int cvro_ggx_sample(const float alpha[2], float eta, const float wi[3],
                    const float u[3], float wo[3], float* weight, int* n_used);
float cvro_fresnel_dielectric(float eta, float ndotwi, float* ndotwt);
float cvro_ggx_g1(const float alpha[2], const float v[3], const float m[3]);
float cvro_density_lookup(const cvro_scene* sc, const float p01[3]);
void cvro_albedo_lookup(const cvro_scene* sc, const float p01[3], float rgb[3]);
float cvro_woodcock(const cvro_scene* sc, const float o[3], const float d[3], float max_t,
                    int32_t seed, int* scattered);
void cvro_camera_ray(const cvro_camera* cam, uint32_t image_id, float u0, float u1,
                     float o[3], float d[3]);
uint32_t cvro_utilhash(uint32_t a);
uint32_t cvro_morton3d(float x, float y, float z);

/* Tile table (Config.h:61-72, CudaVolPath.cpp:12-29).  origins = n_tiles_x*n_tiles_y
 * pairs. */
void cvro_tile_table(uint32_t res_x, uint32_t res_y, uint32_t ntx, uint32_t nty,
                     uint32_t tile_dim[2], uint32_t* origins);

/* Default camera constants (Camera.h:25-42,63-71; CudaVolPath.cpp:66-85). */
void cvro_default_camera(uint32_t res_x, uint32_t res_y, float fov_x,
                         float inv_view[12], float raster_to_view[2]);

#ifdef __cplusplus
}
#endif
#endif
