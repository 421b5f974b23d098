/*
 * oracle/rt_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's CPU path-tracing path (FP64, scalar branches), used as the
 * checker for the CUDA kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (librt_b200.so, the raytracer host
 * program) never links or calls it.
 *
 * Parity status: PINNED against the reference itself — oracle/_ref/libref_harness.so is the
 * unmodified reference compiled from /root/reference/src, and tests/test_oracle_vs_ref.py requires
 * this restatement (mt19937 mode, rejection samplers) to reproduce its closest hits, camera rays and
 * full renders bit for bit; the committed fixtures under tests/golden/ carry those reference outputs
 * to machines where /root/reference does not exist.  The reference ships no tests or golden vectors
 * of its own; its only known-answer data, the sphere-UV table in Sphere.cpp:129-134,
 * is checked as well.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include "../include/rt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Random source. */
enum {
  ORA_RNG_MT19937 = 0, /* std::mt19937 + generate_canonical<double,53>, one sequential stream
                          (Utility.hpp:16-37): reproduces the reference bit for bit */
  ORA_RNG_PHILOX = 1   /* Philox4x32-10 keyed by (seed; pixel, sample, bounce, stream): the stream the
                          CUDA kernels draw from (24-bit uniforms), still FP64 arithmetic */
};

/* Direction samplers. */
enum {
  ORA_SAMPLER_REJECTION = 0, /* CPU reference: Vec3Utility.hpp:41-62 */
  ORA_SAMPLER_POLAR = 1      /* reference CUDA path: Vec3Utility.cuh:57-70 (what the kernels use) */
};

/* Philox stream ids (counter word 3 = stream << 16 | block). */
enum { ORA_STREAM_CAMERA = 0, ORA_STREAM_SHADE = 1, ORA_STREAM_MEDIUM0 = 2 };

typedef struct ora_scene ora_scene;

ora_scene *ora_scene_create(const rt_scene_desc *desc);
void ora_scene_destroy(ora_scene *s);

/* Camera::initialize (Camera.cpp:31-73). */
void ora_camera_init(const rt_camera_config *cfg, rt_camera *out);

/* Camera::get_ray for all pixels, stratum (s_i,s_j), row-major; mt19937 mode draws sequentially in
 * pixel order after seeding, Philox mode keys every pixel separately. */
void ora_primary_rays(const rt_camera_config *cfg, int rng_kind, int sampler, uint64_t seed, int s_i, int s_j,
                      rt_ray *out);

/* Closest hit.  use_bvh = 0: linear scan in world order (HittableList.cpp:26-42); 1: median-split
 * BVH over the same objects with the list's tie rule.  Media draw from the given rng kind
 * (mt19937: one stream seeded with `seed` for the whole batch; Philox: keyed by the ray's rng_*). */
void ora_trace(ora_scene *s, const rt_ray *rays, int64_t n, int use_bvh, int rng_kind, uint64_t seed,
               rt_hit *hits);

typedef struct ora_counters {
  uint64_t paths;
  uint64_t segments;    /* world.hit calls from ray_color */
  uint64_t node_tests;  /* AABB tests (BVH mode) */
  uint64_t sphere_tests;
  uint64_t quad_tests;
  uint64_t rng_draws;
} ora_counters;

/* Rows [row0,row1) as StaticCamera::render_cpu's serial loop renders them (StaticCamera.cpp:101-131):
 * linear RGB doubles, (row1-row0)*W*3.  single_stratum >= 0 renders one un-normalised dynamic-mode
 * frame sample instead (DynamicCamera.cpp:103-171).  Returns wall seconds. */
double ora_render(ora_scene *s, const rt_camera_config *cfg, int rng_kind, int sampler, uint64_t seed,
                  int use_bvh, int row0, int row1, int single_stratum, double *out, ora_counters *counters);

/* Log every segment the integrator traces during ora_render / ora_trace into buf (up to cap rays);
 * pass NULL to stop.  Gives tests realistic secondary-ray sets. */
void ora_ray_log(rt_ray *buf, int64_t cap);
int64_t ora_ray_log_count(void);

/* to_byte (ColorUtility.hpp:18-23). */
int ora_to_byte(double v);

/* Sphere::hit's UV mapping for a unit outward normal (Sphere.cpp:136-140); KAT hook. */
void ora_sphere_uv(const double n[3], double *u, double *v);

/* Raw generators, for pinning against the reference's engine. */
typedef struct ora_mt19937 {
  uint32_t mt[624];
  int idx;
} ora_mt19937;
void ora_mt_seed(ora_mt19937 *g, uint32_t seed);
uint32_t ora_mt_next(ora_mt19937 *g);
double ora_mt_canonical(ora_mt19937 *g);
int ora_mt_uniform_int(ora_mt19937 *g, int lo, int hi);
void ora_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif
