// rt_internal.h — host-side structures shared by the translation units of librt_b200.so and the
// launch wrappers of the kernels (rt_kernels.cu, rt_exact.cu).
#pragma once

#include "../../include/rt_b200.h"
#include "rt_bigvec.h"
#include "rt_bvh.h"
#include "rt_device.h"
#include "rt_exact.h"

#include <cuda_runtime.h>
#include <string>
#include <vector>

// One wavefront pass: `n_paths` = owned pixels x samples of this pass.  Path p covers owned pixel
// p % n_owned and sample first_sample + p / n_owned of a sqrt_spp x sqrt_spp stratification.
struct PassParams {
  DCamera cam;
  DFilmMap map;
  int n_owned;      // owned pixels
  int n_paths;      // paths of this pass (of this sub-pass when the pass is split, rt_api.cu render_pass)
  int path_base;    // the launch sequence covers paths path_base ... path_base + n_paths - 1 (queue 0 slot q = path path_base + q)
  int first_sample; // linear stratum index of the pass's first sample
  int n_samples;    // samples in this pass
  int sqrt_spp;
  float recip_sqrt_spp;
  int max_depth;
  uint64_t seed;
  float4 *film_direct; // one-sample passes: the film, written by the kernel that ends a path; else null
  PathMap paths;        // path id -> (sample, pixel) (rt_device.h)
  FastDiv div_sqrt_spp; // stratum -> (s_i, s_j)
};

// Philox key and film index of a path at a bounce (shared by the render kernels and the parity audit).
RT_HD void path_to_key(const PassParams &pp, int path, int bounce, RayKey &key, int &owned_pixel, int &row, int &col) {
  uint32_t s_local, k;
  path_to_pixel(pp.paths, (uint32_t)path, s_local, k, row, col);
  key.seed = pp.seed;
  key.pixel = (uint32_t)row * (uint32_t)pp.map.width + (uint32_t)col;
  key.sample = (uint32_t)pp.first_sample + s_local;
  key.bounce = (uint32_t)bounce;
  owned_pixel = (int)k;
}
RT_HD void path_to_key(const PassParams &pp, int path, int bounce, RayKey &key, int &owned_pixel) {
  int row, col;
  path_to_key(pp, path, bounce, key, owned_pixel, row, col);
}

// Parity audit (rt_context_set_audit): per queue slot, the FP64 parity traversal's answer for the ray the FP32
// extend kernel is about to trace, and the tallies of the comparison.
enum { RT_AUDIT_SEGMENTS = 0, RT_AUDIT_MISMATCH = 1, RT_AUDIT_PRIMARY = 2, RT_AUDIT_PRIMARY_MISMATCH = 3,
       RT_AUDIT_HIT_MISS = 4, RT_AUDIT_T_ABOVE_1E4 = 5, RT_AUDIT_MAX_REL_T = 6, RT_AUDIT_SAMPLES = 7, RT_AUDIT_WORDS = 8 };
#define RT_AUDIT_MAX_SAMPLES 4096

// Per-context wavefront storage (sized for the largest pass so far).
struct WaveBuffers {
  float4 *ray_a[2] = {nullptr, nullptr}; // origin.xyz, time
  float4 *ray_b[2] = {nullptr, nullptr}; // direction.xyz, path id (int bits)
  float2 *hit[2] = {nullptr, nullptr};   // in: (-, skip primitive)  out: (t, primitive) for the same queue slot
  float4 *thr[2] = {nullptr, nullptr};   // path throughput of the ray in the same queue slot
  float4 *radiance = nullptr;            // per path: final contribution
  unsigned int *counts = nullptr;        // queue length per bounce (max_depth + 2 entries)
  unsigned long long *stats = nullptr;   // [0] segments of the extend launches  [1] node visits  [2] primitive tests
                                         // ([1], [2]: only counted by the instrumented kernels, rt_context_set_stats)
                                         // [3] segments of the tail kernel
  // parity audit (allocated on first use)
  int2 *audit_prim = nullptr;            // per queue slot: leaf-order primitive of the FP64 traversal (-1 = miss), skipped primitive
  double *audit_t = nullptr;             // per queue slot: its t
  unsigned long long *audit_stats = nullptr; // RT_AUDIT_WORDS tallies
  rt_audit_sample *audit_samples = nullptr;  // first RT_AUDIT_MAX_SAMPLES mismatching segments
  size_t capacity_audit = 0;
  size_t last_counts = 0; // queue-length entries the most recent pass used
  size_t last_splits = 1; // sub-passes of the most recent pass: each has its own block of count / cursor words ...
  size_t last_stride = 0; // ... this many words apart
  size_t capacity_paths = 0;
  size_t capacity_counts = 0;
};

struct StageTimer {
  bool enabled = false;
  std::vector<cudaEvent_t> pool;               // reusable events
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> spans; // (stage, (begin, end))
  size_t next = 0;
};

struct rt_context {
  int device = 0;
  int sm_count = 0;
  // Bounces run as separate extend / shade launches before the tail kernel takes over; < 0 = by scene size
  // (rt_api.cu, wave_depth).  RT_WAVE_BOUNCES overrides.
  int wave_bounces = -1;
  // device staging area of the host-pointer convenience calls (rt_film_resolve_rgb8, rt_film_read_rgb): kept
  // across calls so that a frame loop does not pay a cudaMalloc / cudaFree pair per frame
  void *scratch = nullptr;
  size_t scratch_bytes = 0;
  int tail_span = 1 << 20; // bounces covered by one tail launch (measured: one launch for the whole tail is
                           // fastest, even at depth 50; shorter spans chain launches through the queues)
  int64_t pass_paths = (int64_t)16 << 20; // static renders: paths per wavefront pass (queue storage ~110 B per
                                          // path; measured 4 M / 8 M / 16 M / 32 M: 16 M is fastest on C1 and C3)
  int64_t pass_paths_long = (int64_t)64 << 20; // long static renders (rt_api.cu, wave_depth): with six wavefront bounces
                                               // larger passes keep paying - final scene 4K, Mpath-samples/s for 8 / 16 /
                                               // 32 / 64 / 128 M paths per pass: 1903 / 2033 / 2102 / 2139 / 2155 (7 GB of
                                               // queues at 64 M); RT_PASS_PATHS sets both
  // RT_FUSED_GENERATE=1: the first extend launch derives the camera rays itself and the first shade launch
  // re-derives them (no k_generate launch, queue 0 never written: 64 B per path of queue memory and traffic less).
  // Measured on B200 (profiles/r02_experiments.md): the saved launch (0.025 ms on C2) and traffic are paid back by
  // the camera-ray arithmetic done twice at the extend kernel's lane utilisation (+0.014 ms extend, +0.013 ms
  // shade), so the separate k_generate stays the default.
  bool fused_generate = false;
  bool use_graph = true;            // render passes are submitted as one CUDA-graph launch (rt_context_set_graph, RT_GRAPH=0)
  cudaGraphExec_t graph_exec = nullptr; // updated in place pass after pass
  bool audit = false; // every extend launch is checked against the FP64 parity traversal (all-wavefront schedule)
  bool stats = false; // instrumented extend / tail kernels count node visits and primitive tests
  bool defer_accumulate = true;    // multi-sample passes leave their in-order sum to the film's next reader (RT_DEFER_ACCUMULATE=0)
  rt_film *pending_film = nullptr; // the film whose pending sum lives in wave.radiance (at most one per context)
  cudaStream_t stream = nullptr;
  // Small passes (a 1-spp frame) are cut into `split` sub-passes over disjoint path ranges, each with its own queues,
  // submitted on forked streams (parallel branches of the pass's CUDA graph): every persistent kernel ends with its
  // slowest rays on a mostly idle GPU, and the other sub-pass's kernels fill those SMs (rt_api.cu render_pass).
  int split = 1;                          // RT_SPLIT (measured: no gain at 2, slower at 3-4; profiles/r02_experiments.md)
  int64_t split_max_paths = (int64_t)4 << 20; // larger passes amortise their kernel ends: not split
  cudaStream_t side_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t fork_event = nullptr, join_event[3] = {nullptr, nullptr, nullptr};
  cudaStream_t pass_stream = nullptr;     // the stream the render-kernel wrappers launch on (stream or a side stream)
  WaveBuffers wave;
  rt_counters counters{};
  StageTimer timer;
};

struct rt_scene {
  rt_context *ctx = nullptr;
  DScene d{};
  ExactScene ex{};
  // device allocations (owned)
  float4 *nodes = nullptr, *prims = nullptr, *bprims = nullptr, *mats = nullptr, *lights = nullptr,
         *perlin_grad = nullptr;
  unsigned char *perlin_perm = nullptr;
  uint32_t *texels = nullptr;
  PrimExact *ex_prims = nullptr, *ex_bprims = nullptr;
  XformOpExact *ex_ops = nullptr;
  int *ex_chain_first = nullptr, *ex_chain_count = nullptr;
  int *leaf_object = nullptr, *leaf_id = nullptr; // per leaf-order primitive: object / unified id
  rt_scene_info info{};
  int n_leaf = 0;
  // kept for rt_scene_update_spheres (animated scenes): the instance chains and materials the records are
  // baked with, the current spheres, where each sphere's leaf is, and the refit links (made at the first update)
  std::vector<rt_xform> h_xforms;
  std::vector<rt_xform_op> h_xform_ops;
  rtflat::BigVec<rt_sphere> h_spheres;
  rtflat::BigVec<rt_quad> h_quads;
  std::vector<int> sphere_leaf, quad_leaf; // -1: boundary primitive of a medium (not a leaf of its own)
  rtflat::BigVec<float4> h_mats;
  rtflat::BigVec<PrimExact> h_ex_prims; // FP64 parity records (description order) until their first use (rt_scene_ensure_exact)
  rtflat::BigVec<uint32_t> h_order;  // leaf j holds record h_order[j]
  int *leaf_up = nullptr;            // device: per leaf, parent node * 4 + slot
  unsigned int *arrivals = nullptr;  // device: per node refit counter
};

// The per-path radiance of a multi-sample pass that has not been summed into the film yet: k_accumulate's work,
// handed to whichever kernel reads the film next (the present kernel sums while it tone-maps).
struct PendingSum {
  const float4 *radiance = nullptr; // radiance[s * n_owned + path], s < n_samples
  int n_samples = 0;                // 0: nothing pending
  int n_owned = 0;
  int tiled = 0, blocks_x = 1, width = 1; // PathMap: how film index k maps to the path index of its pixel
};
// film index (row-major over the owned scanlines) -> path index of that pixel within one sample (inverse of path_to_pixel)
RT_HD uint32_t film_index_to_path(const PendingSum &p, uint32_t k) {
  if (!p.tiled)
    return k;
  uint32_t local_row = k / (uint32_t)p.width, col = k - local_row * (uint32_t)p.width;
  uint32_t block = (local_row >> 2) * (uint32_t)p.blocks_x + (col >> 3);
  return (block << 5) | ((local_row & 3u) << 3) | (col & 7u);
}

struct rt_film {
  rt_context *ctx = nullptr;
  DFilmMap map{};
  int64_t n_owned = 0;
  float4 *accum = nullptr;
  bool owns_accum = false;
  int64_t samples = 0;
  // a multi-sample pass whose in-order sum into `accum` is still to be done (only films that own their buffer defer
  // it; rt_api.cu film_flush)
  PendingSum pending{};
  PassParams pending_pass{};
};

// ---- error reporting (rt_api.cu) ----
void rt_set_error(const std::string &msg);
int rt_cuda_fail(cudaError_t e, const char *what);
#define RT_CUDA(call)                                                                                        \
  do {                                                                                                       \
    cudaError_t e_ = (call);                                                                                 \
    if (e_ != cudaSuccess)                                                                                   \
      return rt_cuda_fail(e_, #call);                                                                        \
  } while (0)

// ---- scene construction (rt_scene.cu) ----
int rt_scene_build(rt_context *ctx, const rt_scene_desc *desc, rt_scene *scene);
void rt_scene_release(rt_scene *scene);
int rt_scene_ensure_exact(rt_scene *scene);
int rt_scene_update_spheres_impl(rt_scene *scene, int first, int count, const rt_sphere *spheres);
int rt_scene_update_quads_impl(rt_scene *scene, int first, int count, const rt_quad *quads);

// ---- kernel launch wrappers (rt_kernels.cu); all asynchronous on `stream` ----
struct LaunchShape {
  int blocks, threads;
};
LaunchShape rt_persistent_shape(const rt_context *ctx, int threads, int blocks_per_sm);

// LBVH build
void launch_morton(cudaStream_t s, const BuildBox *boxes, int n, const float *scene_lo, const float *scene_inv,
                   uint64_t *codes, uint32_t *index);
int sort_pairs(cudaStream_t s, uint64_t *keys_in, uint64_t *keys_out, uint32_t *vals_in, uint32_t *vals_out, int n);
void launch_gather_boxes(cudaStream_t s, const BuildBox *in, const uint32_t *index, BuildBox *out, int n);
int ploc_build(cudaStream_t s, const BuildBox *leaf_boxes, int n, BinTree t, PlocCluster *clusters[2], int *nearest,
               unsigned long long *packed, int *rounds_out);
int tree_area(cudaStream_t s, const BuildBox *box, int n_internal, double *d_sum, double *out);
void launch_hierarchy(cudaStream_t s, const uint64_t *codes, BinTree t);
void launch_refit(cudaStream_t s, BinTree t, const BuildBox *leaf_boxes);
void launch_collapse(cudaStream_t s, BinTree t, const BuildBox *leaf_boxes, float4 *nodes, const CollapseItem *items,
                     int n_items, CollapseItem *next, int *next_count, int *wide_count);
void launch_gather_records(cudaStream_t s, const void *in, const uint32_t *index, void *out, int n, int bytes_per);

// wavefront render
void launch_generate(const rt_context *ctx, const PassParams &pp, WaveBuffers &w);
void launch_extend(const rt_context *ctx, const DScene &sc, const PassParams &pp, WaveBuffers &w, int bounce, bool gen);
void launch_shade(const rt_context *ctx, const DScene &sc, const PassParams &pp, WaveBuffers &w, int bounce, bool gen);
void launch_tail(const rt_context *ctx, const DScene &sc, const PassParams &pp, WaveBuffers &w, int first_bounce,
                 int end_bounce, int buffer);
void launch_leaf_links(cudaStream_t s, const float4 *nodes, int n_nodes, int *leaf_up);
void launch_update_leaves(cudaStream_t s, const float4 *records, const PrimExact *exact, const BuildBox *boxes,
                          const int *leaf, int count, const int *leaf_up, float4 *prims, PrimExact *ex_prims, float4 *nodes);
void launch_refit_wide(cudaStream_t s, float4 *nodes, const int *leaf_up, unsigned int *arrivals, int n_leaf);
void launch_accumulate(const rt_context *ctx, const PassParams &pp, WaveBuffers &w, float4 *film);
void launch_resolve_rgb8(cudaStream_t s, const float4 *film, int64_t n, double scale, uint8_t *out);
void launch_resolve_rgb(cudaStream_t s, const float4 *film, int64_t n, double scale, float *out);
void launch_scatter_gathered(cudaStream_t s, int width, int height, int n_ranks, int tile_rows, const void *gathered,
                             void *full, int bytes_per_pixel); // 16: float4 sums, 3: RGB8

// displayed frames (rt_frame.cu)
void launch_present_rgb8(const rt_context *ctx, cudaStream_t stream, const float4 *accum, int64_t n_owned, int width,
                         int tile_rows, int rank, int n_ranks, double scale, uint8_t *frame, unsigned int *blocks_done,
                         uint32_t *flag, uint32_t ticket, const uint32_t *consumed, uint32_t *error,
                         float4 *accum_rw = nullptr, const PendingSum &pending = PendingSum());
int rt_film_flush(rt_film *film); // completes a deferred in-order sum (rt_api.cu)

// parity audit (rt_exact.cu)
void launch_audit_trace(const rt_context *ctx, const ExactScene &sc, const PassParams &pp, WaveBuffers &w, int bounce,
                        int has_media);
void launch_audit_compare(const rt_context *ctx, const PassParams &pp, WaveBuffers &w, int bounce, const int *leaf_id);

// parity hook
void launch_trace_fast(const rt_context *ctx, const DScene &sc, const rt_ray *d_rays, int64_t n, uint64_t seed,
                       const int *leaf_object, const int *leaf_id, rt_hit *d_hits);
void launch_trace_exact(cudaStream_t s, const ExactScene &sc, const rt_ray *d_rays, int64_t n, uint64_t seed,
                        rt_hit *d_hits); // rt_exact.cu
