// rt_api.cu — the C ABI of librt_b200.so (include/rt_b200.h): contexts, scenes, films and the host
// side of the wavefront render loop.  There is no CPU fallback anywhere in this library: without a
// CUDA device every entry point that needs one returns RT_ERR_NO_DEVICE.
#include "rt_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

static thread_local std::string g_last_error;

void rt_set_error(const std::string &msg) { g_last_error = msg; }

int rt_cuda_fail(cudaError_t e, const char *what) {
  g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? RT_ERR_NO_DEVICE : RT_ERR_CUDA;
}

namespace {

int invalid(const char *msg) {
  rt_set_error(msg);
  return RT_ERR_INVALID;
}

// The context's staging area, at least `bytes` large.  Callers synchronise the stream before they return, so
// one area serves every call.
int ctx_scratch(rt_context *ctx, size_t bytes, void **out) {
  if (bytes > ctx->scratch_bytes) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->scratch);
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    RT_CUDA(cudaMalloc(&ctx->scratch, bytes));
    ctx->scratch_bytes = bytes;
  }
  *out = ctx->scratch;
  return RT_OK;
}

// Grow the per-context wavefront storage.
int ensure_wave(rt_context *ctx, size_t n_paths, size_t n_counts) {
  WaveBuffers &w = ctx->wave;
  if (n_paths > w.capacity_paths) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 2; k++) {
      cudaFree(w.ray_a[k]);
      cudaFree(w.ray_b[k]);
      cudaFree(w.hit[k]);
      cudaFree(w.thr[k]);
      w.ray_a[k] = w.ray_b[k] = w.thr[k] = nullptr;
      w.hit[k] = nullptr;
    }
    cudaFree(w.radiance);
    w.radiance = nullptr;
    w.capacity_paths = 0;
    for (int k = 0; k < 2; k++) {
      RT_CUDA(cudaMalloc((void **)&w.ray_a[k], n_paths * sizeof(float4)));
      RT_CUDA(cudaMalloc((void **)&w.ray_b[k], n_paths * sizeof(float4)));
      RT_CUDA(cudaMalloc((void **)&w.hit[k], n_paths * sizeof(float2)));
      RT_CUDA(cudaMalloc((void **)&w.thr[k], n_paths * sizeof(float4)));
    }
    RT_CUDA(cudaMalloc((void **)&w.radiance, n_paths * sizeof(float4)));
    w.capacity_paths = n_paths;
  }
  if (n_counts > w.capacity_counts) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(w.counts);
    w.counts = nullptr;
    RT_CUDA(cudaMalloc((void **)&w.counts, n_counts * sizeof(unsigned int)));
    w.capacity_counts = n_counts;
  }
  if (!w.stats) {
    RT_CUDA(cudaMalloc((void **)&w.stats, 8 * sizeof(unsigned long long)));
    RT_CUDA(cudaMemsetAsync(w.stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
  }
  return RT_OK;
}

// Storage of the parity audit, sized like the queues.
int ensure_audit(rt_context *ctx, size_t n_paths) {
  WaveBuffers &w = ctx->wave;
  if (n_paths > w.capacity_audit) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(w.audit_prim);
    cudaFree(w.audit_t);
    w.audit_prim = nullptr;
    w.audit_t = nullptr;
    w.capacity_audit = 0;
    RT_CUDA(cudaMalloc((void **)&w.audit_prim, n_paths * sizeof(int2)));
    RT_CUDA(cudaMalloc((void **)&w.audit_t, n_paths * sizeof(double)));
    w.capacity_audit = n_paths;
  }
  if (!w.audit_stats) {
    RT_CUDA(cudaMalloc((void **)&w.audit_stats, RT_AUDIT_WORDS * sizeof(unsigned long long)));
    RT_CUDA(cudaMemsetAsync(w.audit_stats, 0, RT_AUDIT_WORDS * sizeof(unsigned long long), ctx->stream));
    RT_CUDA(cudaMalloc((void **)&w.audit_samples, RT_AUDIT_MAX_SAMPLES * sizeof(rt_audit_sample)));
  }
  return RT_OK;
}

DCamera to_device_camera(const rt_camera *c) {
  DCamera d{};
  for (int a = 0; a < 3; a++) {
    d.center[a] = (float)c->center[a];
    d.p00c[a] = (float)(c->pixel00_loc[a] - c->center[a]);
    d.du[a] = (float)c->pixel_delta_u[a];
    d.dv[a] = (float)c->pixel_delta_v[a];
    d.disk_u[a] = (float)c->defocus_disk_u[a];
    d.disk_v[a] = (float)c->defocus_disk_v[a];
  }
  d.defocus = c->defocus_angle > 0;
  d.width = c->image_width;
  d.height = c->image_height;
  return d;
}

cudaEvent_t timer_event(StageTimer &t) {
  if (t.next == t.pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    t.pool.push_back(e);
  }
  return t.pool[t.next++];
}

// Brackets one launch with events when stage timing is on.
struct StageSpan {
  rt_context *ctx;
  int stage;
  cudaEvent_t begin = nullptr;
  StageSpan(rt_context *c, int s) : ctx(c), stage(s) {
    if (ctx->timer.enabled) {
      begin = timer_event(ctx->timer);
      cudaEventRecord(begin, ctx->stream);
    }
  }
  ~StageSpan() {
    if (begin) {
      cudaEvent_t end = timer_event(ctx->timer);
      cudaEventRecord(end, ctx->stream);
      ctx->timer.spans.push_back({stage, {begin, end}});
    }
  }
};

// How many bounces run as wavefront launches before k_tail takes the rest.  Every wavefront launch ends with
// its slowest rays, the tail kernel pays that once, and the larger the scene the longer the slowest rays: measured
// on 1080p frames of the spheres scene (ms for 1 / 2 / 3 wavefront bounces): 485 primitives 0.681 / 0.657 / 0.657,
// 4 k 0.896 / 0.865 / 0.873, 16 k 0.968 / 0.965 / 1.020, 65 k 1.33 / 1.39 / 1.60, 10^6 2.15 / 2.70 / 3.50; with the
// kernels of the end of round 2: 485 primitives 0.643 / 0.619 / 0.619 / 0.624 (4), 14 k 0.915 / 0.902 / 0.949,
// 10^6 1.80 / 2.35 / 3.01, Cornell box 0.510 / 0.490 / 0.488 / 0.489.
// `long_render`: the call is a static render of many samples (>= 32 M paths over the whole image, whatever the number
// of GPUs: the choice must not depend on the rank count, or the image would).  Its passes are 16 M paths each, the
// launch ends weigh little, and the wavefront kernels - better lane utilisation than the tail kernel - pay for more
// bounces: final scene 3840x2160, 256 spp, Mpath-samples/s for 1 / 2 / 3 / 4 / 6 / 8 / 12 wavefront bounces: 1770 /
// 1884 / 1949 / 1991 / 2032 / 2037 / 2032; Cornell + smoke 1080p 5280 (2) / 5317 / 5366 (4, 6) / 5301 (8); 10^6 spheres
// 64 spp 99.5 / 97.8 / 100.8 / 104.6 ms.
static int wave_depth(const rt_context *ctx, const rt_scene *scene, bool long_render) {
  if (ctx->audit) // the audit brackets extend launches; the tail kernel traces inside one launch
    return 1 << 20;
  if (ctx->wave_bounces >= 0)
    return ctx->wave_bounces;
  // (Round 1 also sent every scene with participating media above 2,048 primitives to depth 1; that measurement was
  // distorted by rays with a zero direction component walking the whole tree - rt_device.h, RayTrav.  Re-measured on
  // the final scene with its two media: 0.893 / 0.872 / 0.883 / 0.899 ms for 1 / 2 / 3 / 4.)
  if (long_render)
    return scene->n_leaf > 32768 ? 2 : (scene->n_leaf > 2048 ? 6 : 4);
  return scene->n_leaf > 32768 ? 1 : (scene->n_leaf > 2048 ? 2 : 3);
}

} // namespace

// A multi-sample pass on a film that owns its buffer leaves the in-order per-pixel sum of its samples to the next
// reader of the film: rt_film_present folds it into its tone-map kernel (one launch and one pass over the film
// less per displayed frame); every other reader, and the next render pass of the context (which reuses the
// radiance buffer), completes it here with k_accumulate.  Same operands in the same order either way.
int rt_film_flush(rt_film *film) {
  if (!film || !film->pending.n_samples)
    return RT_OK;
  rt_context *ctx = film->ctx;
  launch_accumulate(ctx, film->pending_pass, ctx->wave, film->accum);
  ctx->counters.kernel_launches += 1;
  film->pending = PendingSum();
  if (ctx->pending_film == film)
    ctx->pending_film = nullptr;
  RT_CUDA(cudaGetLastError());
  return RT_OK;
}

namespace {

// One wavefront pass over `n_samples` strata starting at linear stratum `first_sample`.
int render_pass(rt_scene *scene, const rt_camera *camera, rt_film *film, int first_sample, int n_samples,
                int sqrt_spp, int max_depth, uint64_t seed, bool long_render = false) {
  rt_context *ctx = scene->ctx;
  PassParams pp{};
  pp.cam = to_device_camera(camera);
  pp.map = film->map;
  pp.n_owned = (int)film->n_owned;
  pp.n_paths = (int)(film->n_owned * n_samples);
  pp.first_sample = first_sample;
  pp.n_samples = n_samples;
  pp.sqrt_spp = sqrt_spp;
  pp.recip_sqrt_spp = (float)(1.0 / sqrt_spp);
  pp.max_depth = max_depth;
  pp.seed = seed;
  pp.film_direct = n_samples == 1 ? film->accum : nullptr;
  pp.div_sqrt_spp = fastdiv_make((uint32_t)std::max(sqrt_spp, 1));
  {
    static const bool rows_only = getenv("RT_PATH_ORDER") && std::string(getenv("RT_PATH_ORDER")) == "rows";
    pp.paths = pathmap_make(pp.map, film->n_owned, !rows_only);
  }
  if (pp.n_paths == 0)
    return RT_OK;
  if (ctx->pending_film) { // its radiance lives in the buffer this pass is about to overwrite
    int flushed = rt_film_flush(ctx->pending_film);
    if (flushed != RT_OK)
      return flushed;
  }
  // Sub-passes (rt_internal.h, rt_context::split): a small pass is cut into disjoint path ranges (multiples of a
  // block of 128 paths), each with its own queues and its own block of count / cursor words, launched on forked
  // streams.  The profiling modes bracket or count individual launches on one stream and keep the single sequence.
  const size_t count_words = 3 * ((size_t)max_depth + 2); // queue lengths + fetch cursors of the extend / tail and of the shade launches
  int n_split = 1;
  if (ctx->split > 1 && !ctx->timer.enabled && !ctx->audit && !ctx->stats && pp.n_paths >= 65536 &&
      (int64_t)pp.n_paths <= ctx->split_max_paths)
    n_split = std::min(ctx->split, 4);
  int st = ensure_wave(ctx, (size_t)pp.n_paths, count_words * (size_t)n_split);
  if (st != RT_OK)
    return st;
  DScene sc = scene->d;
  for (int a = 0; a < 3; a++)
    sc.bg[a] = (float)camera->background[a];
  if (ctx->audit && ((st = ensure_audit(ctx, (size_t)pp.n_paths)) != RT_OK || (st = rt_scene_ensure_exact(scene)) != RT_OK))
    return st;
  WaveBuffers &w = ctx->wave;
  // The pass as ONE graph launch (RT_GRAPH=1 / rt_context_set_graph): the launch sequence below is captured from
  // the stream every pass (capturing costs a fraction of launching) and the context's executable graph is updated
  // in place with the new kernel arguments - camera, seed, stratum - so that the driver receives a single launch
  // per pass.  The profiling modes (stage timing, audit) bracket individual launches and keep the direct path.
  const bool as_graph = ctx->use_graph && !ctx->timer.enabled && !ctx->audit;
  if (as_graph)
    RT_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
  RT_CUDA(cudaMemsetAsync(w.counts, 0, count_words * (size_t)n_split * sizeof(unsigned int), ctx->stream));
  // wavefront launches while the path population is large, then one tail kernel that runs whatever is
  // left to completion (rt_kernels.cu, k_tail)
  const int wave_bounces = std::min(max_depth, wave_depth(ctx, scene, long_render));
  // The first extend launch derives the camera rays itself and queue 0 is never written (k_extend<GEN>), unless
  // something else reads queue 0: a tail-only schedule, the parity audit, RT_FUSED_GENERATE=0 (A/B aid).
  const bool fused_generate = wave_bounces >= 1 && !ctx->audit && ctx->fused_generate;
  // the sub-passes: parameters, queue views and streams
  PassParams sub_pp[4];
  WaveBuffers sub_w[4];
  cudaStream_t sub_stream[4];
  {
    const int blocks_total = (pp.n_paths + 127) / 128;
    int first_block = 0;
    for (int h = 0; h < n_split; h++) {
      const int end_block = (int)((int64_t)blocks_total * (h + 1) / n_split);
      const int first = first_block * 128, end = std::min(pp.n_paths, end_block * 128);
      sub_pp[h] = pp;
      sub_pp[h].path_base = first;
      sub_pp[h].n_paths = end - first;
      sub_w[h] = w;
      for (int k = 0; k < 2; k++) {
        sub_w[h].ray_a[k] = w.ray_a[k] + first;
        sub_w[h].ray_b[k] = w.ray_b[k] + first;
        sub_w[h].hit[k] = w.hit[k] + first;
        sub_w[h].thr[k] = w.thr[k] + first;
      }
      sub_w[h].counts = w.counts + count_words * (size_t)h;
      sub_stream[h] = h == 0 ? ctx->stream : ctx->side_stream[h - 1];
      first_block = end_block;
    }
  }
  if (n_split > 1) { // fork: the side streams start behind the memset (and join the capture)
    RT_CUDA(cudaEventRecord(ctx->fork_event, ctx->stream));
    for (int h = 1; h < n_split; h++)
      RT_CUDA(cudaStreamWaitEvent(sub_stream[h], ctx->fork_event, 0));
  }
  // launches are issued stage by stage across the sub-passes, so that without a graph the streams fill evenly
  if (!fused_generate) {
    StageSpan span(ctx, RT_STAGE_GENERATE);
    for (int h = 0; h < n_split; h++) {
      ctx->pass_stream = sub_stream[h];
      launch_generate(ctx, sub_pp[h], sub_w[h]);
    }
  }
  for (int bounce = 0; bounce < wave_bounces; bounce++) {
    const bool gen = fused_generate && bounce == 0;
    if (ctx->audit)
      launch_audit_trace(ctx, scene->ex, pp, w, bounce, sc.n_media > 0);
    {
      StageSpan span(ctx, RT_STAGE_EXTEND);
      for (int h = 0; h < n_split; h++) {
        ctx->pass_stream = sub_stream[h];
        launch_extend(ctx, sc, sub_pp[h], sub_w[h], bounce, gen);
      }
    }
    if (ctx->audit) {
      launch_audit_compare(ctx, pp, w, bounce, scene->leaf_id);
      ctx->counters.kernel_launches += 2;
    }
    {
      StageSpan span(ctx, RT_STAGE_SHADE);
      for (int h = 0; h < n_split; h++) {
        ctx->pass_stream = sub_stream[h];
        launch_shade(ctx, sc, sub_pp[h], sub_w[h], bounce, gen);
      }
    }
  }
  // each tail launch covers at most tail_span bounces and queues its survivors for the next one
  int tail_launches = 0;
  for (int first = wave_bounces, buffer = wave_bounces & 1; first < max_depth; first += ctx->tail_span, buffer ^= 1) {
    StageSpan span(ctx, RT_STAGE_TAIL);
    for (int h = 0; h < n_split; h++) {
      ctx->pass_stream = sub_stream[h];
      launch_tail(ctx, sc, sub_pp[h], sub_w[h], first, std::min(max_depth, first + ctx->tail_span), buffer);
    }
    tail_launches++;
  }
  ctx->pass_stream = ctx->stream;
  if (n_split > 1) { // join: whatever follows on the context stream sees the whole pass
    for (int h = 1; h < n_split; h++) {
      RT_CUDA(cudaEventRecord(ctx->join_event[h - 1], sub_stream[h]));
      RT_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->join_event[h - 1], 0));
    }
  }
  // the in-order sum of a multi-sample pass: deferred to the film's next reader when the film owns its buffer
  // (nobody can look at it behind the library's back) and no per-launch timing wants the launch here
  const bool defer_sum = !pp.film_direct && film->owns_accum && !ctx->timer.enabled && ctx->defer_accumulate;
  if (!pp.film_direct && !defer_sum) {
    StageSpan span(ctx, RT_STAGE_ACCUMULATE);
    launch_accumulate(ctx, pp, w, film->accum);
  }
  if (defer_sum) {
    film->pending.radiance = w.radiance;
    film->pending.n_samples = pp.n_samples;
    film->pending.n_owned = pp.n_owned;
    film->pending.tiled = pp.paths.tiled;
    film->pending.blocks_x = pp.paths.blocks_x;
    film->pending.width = pp.map.width > 0 ? pp.map.width : 1;
    film->pending_pass = pp;
    ctx->pending_film = film;
  }
  ctx->counters.kernel_launches += ((pp.film_direct || defer_sum) ? 0 : 1) +
                                   (uint64_t)n_split * ((fused_generate ? 0 : 1) + 2 * (uint64_t)wave_bounces + (uint64_t)tail_launches);
  ctx->counters.paths += (uint64_t)pp.n_paths;
  w.last_counts = (size_t)max_depth + 1;
  w.last_splits = (size_t)n_split;
  w.last_stride = count_words;
  if (as_graph) {
    cudaGraph_t graph = nullptr;
    RT_CUDA(cudaStreamEndCapture(ctx->stream, &graph));
    bool ready = false;
    if (ctx->graph_exec) { // same topology as the previous pass: only the node parameters change
      cudaGraphExecUpdateResultInfo info;
      ready = cudaGraphExecUpdate(ctx->graph_exec, graph, &info) == cudaSuccess;
      if (!ready) {
        cudaGetLastError();
        cudaGraphExecDestroy(ctx->graph_exec);
        ctx->graph_exec = nullptr;
      }
    }
    if (!ready) {
      cudaError_t e = cudaGraphInstantiate(&ctx->graph_exec, graph, 0);
      if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        return rt_cuda_fail(e, "cudaGraphInstantiate");
      }
      ctx->counters.graph_instantiations += 1;
    }
    cudaGraphDestroy(graph);
    RT_CUDA(cudaGraphLaunch(ctx->graph_exec, ctx->stream));
    ctx->counters.graph_launches += 1;
  }
  RT_CUDA(cudaGetLastError());
  return RT_OK;
}

} // namespace

extern "C" {

int rt_abi_version(void) { return RT_B200_ABI_VERSION; }

const char *rt_last_error(void) { return g_last_error.c_str(); }

int rt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// Camera::initialize (core/camera/Camera.cpp:31-73), same FP64 operation order.
int rt_camera_init(const rt_camera_config *cfg, rt_camera *out) {
  if (!cfg || !out)
    return invalid("rt_camera_init: null argument");
  if (cfg->image_width < 1 || !(cfg->aspect_ratio > 0) || cfg->samples_per_pixel < 1)
    return invalid("rt_camera_init: image_width, aspect_ratio and samples_per_pixel must be positive");
  const double pi = 3.1415926535897932385;
  struct V {
    double x, y, z;
  };
  auto sub = [](V a, V b) { return V{a.x - b.x, a.y - b.y, a.z - b.z}; };
  auto add = [](V a, V b) { return V{a.x + b.x, a.y + b.y, a.z + b.z}; };
  auto scale = [](double t, V a) { return V{t * a.x, t * a.y, t * a.z}; };
  auto divide = [&](V a, double t) { return scale(1 / t, a); };
  auto cross = [](V a, V b) { return V{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; };
  auto unit = [&](V a) {
    double len = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len > 1e-8) {
      double s = 1.0 / len;
      return V{a.x * s, a.y * s, a.z * s};
    }
    return V{1.0, 0.0, 0.0};
  };
  auto load = [](const double *p) { return V{p[0], p[1], p[2]}; };
  auto store = [](double *p, V a) {
    p[0] = a.x;
    p[1] = a.y;
    p[2] = a.z;
  };
  int W = cfg->image_width;
  int H = int(W / cfg->aspect_ratio);
  H = (H < 1) ? 1 : H;
  V center = load(cfg->lookfrom);
  double theta = cfg->vfov * pi / 180.0;
  double h = std::tan(theta / 2);
  double viewport_height = 2 * h * cfg->focus_dist;
  double viewport_width = viewport_height * (double(W) / H);
  V w = unit(sub(load(cfg->lookfrom), load(cfg->lookat)));
  V u = unit(cross(load(cfg->vup), w));
  V v = cross(w, u);
  V viewport_u = scale(viewport_width, u);
  V viewport_v = scale(viewport_height, V{-v.x, -v.y, -v.z});
  V du = divide(viewport_u, W);
  V dv = divide(viewport_v, H);
  V upper_left = sub(sub(sub(center, scale(cfg->focus_dist, w)), divide(viewport_u, 2)), divide(viewport_v, 2));
  V p00 = add(upper_left, scale(0.5, add(du, dv)));
  double defocus_radius = cfg->focus_dist * std::tan((cfg->defocus_angle / 2) * pi / 180.0);
  std::memset(out, 0, sizeof *out);
  out->image_width = W;
  out->image_height = H;
  store(out->center, center);
  store(out->pixel00_loc, p00);
  store(out->pixel_delta_u, du);
  store(out->pixel_delta_v, dv);
  store(out->defocus_disk_u, scale(defocus_radius, u));
  store(out->defocus_disk_v, scale(defocus_radius, v));
  out->defocus_angle = cfg->defocus_angle;
  std::memcpy(out->background, cfg->background, sizeof out->background);
  return RT_OK;
}

int rt_context_create(int device, rt_context **out) {
  if (!out)
    return invalid("rt_context_create: null output");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    rt_set_error("no CUDA device available; this backend has no CPU fallback");
    return RT_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n)
    return invalid("rt_context_create: device index out of range");
  RT_CUDA(cudaSetDevice(device));
  rt_context *ctx = new (std::nothrow) rt_context();
  if (!ctx)
    return invalid("out of host memory");
  ctx->device = device;
  cudaDeviceProp prop;
  RT_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  if (const char *env = std::getenv("RT_WAVE_BOUNCES")) // tuning / A-B aid: bounces run as wavefront launches
    ctx->wave_bounces = std::max(0, std::atoi(env));
  if (const char *env = std::getenv("RT_DEFER_ACCUMULATE"))
    ctx->defer_accumulate = std::atoi(env) != 0;
  if (const char *env = std::getenv("RT_GRAPH"))
    ctx->use_graph = std::atoi(env) != 0;
  if (const char *env = std::getenv("RT_FUSED_GENERATE"))
    ctx->fused_generate = std::atoi(env) != 0;
  if (const char *env = std::getenv("RT_TAIL_SPAN"))
    ctx->tail_span = std::max(1, std::atoi(env));
  if (const char *env = std::getenv("RT_PASS_PATHS"))
    ctx->pass_paths = ctx->pass_paths_long = std::max<int64_t>(1, std::atoll(env));
  if (const char *env = std::getenv("RT_SPLIT")) // sub-passes of a small pass (1 = one launch sequence)
    ctx->split = std::min(4, std::max(1, std::atoi(env)));
  RT_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->pass_stream = ctx->stream;
  for (int k = 0; k < 3; k++) {
    RT_CUDA(cudaStreamCreateWithFlags(&ctx->side_stream[k], cudaStreamNonBlocking));
    RT_CUDA(cudaEventCreateWithFlags(&ctx->join_event[k], cudaEventDisableTiming));
  }
  RT_CUDA(cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
  *out = ctx;
  return RT_OK;
}

void rt_context_destroy(rt_context *ctx) {
  if (!ctx)
    return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  WaveBuffers &w = ctx->wave;
  for (int k = 0; k < 2; k++) {
    cudaFree(w.ray_a[k]);
    cudaFree(w.ray_b[k]);
    cudaFree(w.hit[k]);
    cudaFree(w.thr[k]);
  }
  cudaFree(w.radiance);
  cudaFree(w.counts);
  cudaFree(w.stats);
  cudaFree(w.audit_prim);
  cudaFree(w.audit_t);
  cudaFree(w.audit_stats);
  cudaFree(w.audit_samples);
  cudaFree(ctx->scratch);
  if (ctx->graph_exec)
    cudaGraphExecDestroy(ctx->graph_exec);
  for (cudaEvent_t e : ctx->timer.pool)
    cudaEventDestroy(e);
  for (int k = 0; k < 3; k++) {
    if (ctx->side_stream[k])
      cudaStreamDestroy(ctx->side_stream[k]);
    if (ctx->join_event[k])
      cudaEventDestroy(ctx->join_event[k]);
  }
  if (ctx->fork_event)
    cudaEventDestroy(ctx->fork_event);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

int rt_context_synchronize(rt_context *ctx) {
  if (!ctx)
    return invalid("null context");
  RT_CUDA(cudaSetDevice(ctx->device));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  return RT_OK;
}

uint64_t rt_context_stream(rt_context *ctx) { return ctx ? (uint64_t)(uintptr_t)ctx->stream : 0; }

int rt_scene_create(rt_context *ctx, const rt_scene_desc *desc, rt_scene **out) {
  if (!ctx || !desc || !out)
    return invalid("rt_scene_create: null argument");
  *out = nullptr;
  RT_CUDA(cudaSetDevice(ctx->device));
  rt_scene *sc = new (std::nothrow) rt_scene();
  if (!sc)
    return invalid("out of host memory");
  int st = rt_scene_build(ctx, desc, sc);
  if (st != RT_OK) {
    rt_scene_release(sc);
    delete sc;
    return st;
  }
  *out = sc;
  return RT_OK;
}

void rt_scene_destroy(rt_scene *scene) {
  if (!scene)
    return;
  cudaSetDevice(scene->ctx->device);
  cudaStreamSynchronize(scene->ctx->stream);
  rt_scene_release(scene);
  delete scene;
}

int rt_scene_update_spheres(rt_scene *scene, int first_sphere, int n_spheres, const rt_sphere *spheres) {
  if (!scene)
    return invalid("rt_scene_update_spheres: null scene");
  RT_CUDA(cudaSetDevice(scene->ctx->device));
  return rt_scene_update_spheres_impl(scene, first_sphere, n_spheres, spheres);
}

int rt_scene_update_quads(rt_scene *scene, int first_quad, int n_quads, const rt_quad *quads) {
  if (!scene)
    return invalid("rt_scene_update_quads: null scene");
  RT_CUDA(cudaSetDevice(scene->ctx->device));
  return rt_scene_update_quads_impl(scene, first_quad, n_quads, quads);
}

int rt_scene_get_info(rt_scene *scene, rt_scene_info *out) {
  if (!scene || !out)
    return invalid("rt_scene_get_info: null argument");
  *out = scene->info;
  return RT_OK;
}

int rt_trace_rays(rt_scene *scene, const rt_ray *rays, int64_t n, int mode, uint64_t seed, rt_hit *hits) {
  if (!scene || (n > 0 && (!rays || !hits)) || n < 0)
    return invalid("rt_trace_rays: bad argument");
  if (mode != RT_TRACE_EXACT_F64 && mode != RT_TRACE_FAST_F32)
    return invalid("rt_trace_rays: unknown mode");
  if (n == 0)
    return RT_OK;
  rt_context *ctx = scene->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  if (mode == RT_TRACE_EXACT_F64) {
    int ready = rt_scene_ensure_exact(scene);
    if (ready != RT_OK)
      return ready;
  }
  rt_ray *d_rays = nullptr;
  rt_hit *d_hits = nullptr;
  RT_CUDA(cudaMalloc((void **)&d_rays, (size_t)n * sizeof(rt_ray)));
  cudaError_t e = cudaMalloc((void **)&d_hits, (size_t)n * sizeof(rt_hit));
  if (e != cudaSuccess) {
    cudaFree(d_rays);
    return rt_cuda_fail(e, "cudaMalloc (hits)");
  }
  int st = RT_OK;
  e = cudaMemcpyAsync(d_rays, rays, (size_t)n * sizeof(rt_ray), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    if (mode == RT_TRACE_EXACT_F64)
      launch_trace_exact(ctx->stream, scene->ex, d_rays, n, seed, d_hits);
    else
      launch_trace_fast(ctx, scene->d, d_rays, n, seed, scene->leaf_object, scene->leaf_id, d_hits);
    ctx->counters.kernel_launches += 1;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(hits, d_hits, (size_t)n * sizeof(rt_hit), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess)
    st = rt_cuda_fail(e, "rt_trace_rays");
  cudaFree(d_rays);
  cudaFree(d_hits);
  return st;
}

int64_t rt_film_owned_pixels_for(int width, int height, int rank, int n_ranks, int tile_rows) {
  if (width < 1 || height < 1 || n_ranks < 1 || rank < 0 || rank >= n_ranks || tile_rows < 1)
    return -1;
  return (int64_t)owned_rows(height, rank, n_ranks, tile_rows) * width;
}

int rt_film_create(rt_context *ctx, int width, int height, int rank, int n_ranks, int tile_rows, void *external_accum,
                   rt_film **out) {
  if (!ctx || !out)
    return invalid("rt_film_create: null argument");
  *out = nullptr;
  int64_t owned = rt_film_owned_pixels_for(width, height, rank, n_ranks, tile_rows);
  if (owned < 0)
    return invalid("rt_film_create: bad geometry");
  if ((int64_t)width * height > (int64_t)1 << 30)
    return invalid("rt_film_create: image too large");
  RT_CUDA(cudaSetDevice(ctx->device));
  rt_film *f = new (std::nothrow) rt_film();
  if (!f)
    return invalid("out of host memory");
  f->ctx = ctx;
  f->map.width = width;
  f->map.height = height;
  f->map.rank = rank;
  f->map.n_ranks = n_ranks;
  f->map.tile_rows = tile_rows;
  f->n_owned = owned;
  if (external_accum) {
    f->accum = (float4 *)external_accum;
    f->owns_accum = false;
  } else {
    cudaError_t e = cudaMalloc((void **)&f->accum, std::max<int64_t>(owned, 1) * sizeof(float4));
    if (e != cudaSuccess) {
      delete f;
      return rt_cuda_fail(e, "cudaMalloc (film)");
    }
    f->owns_accum = true;
  }
  *out = f;
  return rt_film_clear(f);
}

void rt_film_destroy(rt_film *film) {
  if (!film)
    return;
  if (film->ctx->pending_film == film)
    film->ctx->pending_film = nullptr; // an unsummed pass dies with its film
  cudaSetDevice(film->ctx->device);
  cudaStreamSynchronize(film->ctx->stream);
  if (film->owns_accum)
    cudaFree(film->accum);
  delete film;
}

int rt_film_clear(rt_film *film) {
  if (!film)
    return invalid("null film");
  RT_CUDA(cudaSetDevice(film->ctx->device));
  film->pending = PendingSum(); // cleared before it was ever summed
  if (film->ctx->pending_film == film)
    film->ctx->pending_film = nullptr;
  if (film->n_owned > 0)
    RT_CUDA(cudaMemsetAsync(film->accum, 0, (size_t)film->n_owned * sizeof(float4), film->ctx->stream));
  film->samples = 0;
  return RT_OK;
}

int64_t rt_film_owned_pixels(const rt_film *film) { return film ? film->n_owned : -1; }
uint64_t rt_film_device_ptr(rt_film *film) {
  if (!film)
    return 0;
  cudaSetDevice(film->ctx->device);
  rt_film_flush(film); // whoever asks for the raw sums gets complete ones (stream-ordered)
  return (uint64_t)(uintptr_t)film->accum;
}
int64_t rt_film_samples(const rt_film *film) { return film ? film->samples : -1; }

static int check_render_args(rt_scene *scene, const rt_camera *camera, rt_film *film, int sqrt_spp, int max_depth) {
  if (!scene || !camera || !film)
    return invalid("render: null argument");
  if (scene->ctx != film->ctx)
    return invalid("render: scene and film belong to different contexts");
  if (camera->image_width != film->map.width || camera->image_height != film->map.height)
    return invalid("render: camera and film sizes differ");
  if (sqrt_spp < 1 || max_depth < 1 || max_depth > 4096)
    return invalid("render: sqrt_spp and max_depth must be positive");
  return RT_OK;
}

int rt_render_accumulate(rt_scene *scene, const rt_camera *camera, rt_film *film, int s_i, int s_j, int sqrt_spp,
                         int max_depth, uint64_t seed) {
  int st = check_render_args(scene, camera, film, sqrt_spp, max_depth);
  if (st != RT_OK)
    return st;
  if (s_i < 0 || s_i >= sqrt_spp || s_j < 0 || s_j >= sqrt_spp)
    return invalid("rt_render_accumulate: stratum out of range");
  RT_CUDA(cudaSetDevice(scene->ctx->device));
  st = render_pass(scene, camera, film, s_j * sqrt_spp + s_i, 1, sqrt_spp, max_depth, seed);
  if (st == RT_OK)
    film->samples += 1;
  return st;
}

// Strata [first, first + count) in passes of enough paths to fill the GPU several times over, bounded
// so that the queues stay modest.
static int render_strata(rt_scene *scene, const rt_camera *camera, rt_film *film, int first, int count, int sqrt_spp,
                         int max_depth, uint64_t seed) {
  // the schedule of a long static render (wave_depth): decided from the whole image, not from this rank's share
  const bool long_render = (int64_t)count * film->map.width * film->map.height >= ((int64_t)32 << 20);
  const int64_t target_paths = long_render ? scene->ctx->pass_paths_long : scene->ctx->pass_paths;
  int per_pass = (int)std::max<int64_t>(1, std::min<int64_t>(count, target_paths / std::max<int64_t>(film->n_owned, 1)));
  for (int s = first; s < first + count; s += per_pass) {
    int n = std::min(per_pass, first + count - s);
    int st = render_pass(scene, camera, film, s, n, sqrt_spp, max_depth, seed, long_render);
    if (st != RT_OK)
      return st;
    film->samples += n;
  }
  return RT_OK;
}

int rt_render_strata(rt_scene *scene, const rt_camera *camera, rt_film *film, int first_stratum, int n_strata,
                     int sqrt_spp, int max_depth, uint64_t seed) {
  int st = check_render_args(scene, camera, film, sqrt_spp, max_depth);
  if (st != RT_OK)
    return st;
  if (first_stratum < 0 || n_strata < 1 || first_stratum + n_strata > sqrt_spp * sqrt_spp)
    return invalid("rt_render_strata: strata out of range");
  RT_CUDA(cudaSetDevice(scene->ctx->device));
  return render_strata(scene, camera, film, first_stratum, n_strata, sqrt_spp, max_depth, seed);
}

int rt_render_static(rt_scene *scene, const rt_camera *camera, rt_film *film, int sqrt_spp, int max_depth,
                     uint64_t seed) {
  int st = check_render_args(scene, camera, film, sqrt_spp, max_depth);
  if (st != RT_OK)
    return st;
  RT_CUDA(cudaSetDevice(scene->ctx->device));
  if ((st = rt_film_clear(film)) != RT_OK)
    return st;
  return render_strata(scene, camera, film, 0, sqrt_spp * sqrt_spp, sqrt_spp, max_depth, seed);
}

int rt_film_read_rgb(rt_film *film, double scale, float *host_rgb) {
  if (!film || !host_rgb)
    return invalid("rt_film_read_rgb: null argument");
  if (film->n_owned == 0)
    return RT_OK;
  rt_context *ctx = film->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  if (int flushed = rt_film_flush(film))
    return flushed;
  void *scratch = nullptr;
  int st = ctx_scratch(ctx, (size_t)film->n_owned * 3 * sizeof(float), &scratch);
  if (st != RT_OK)
    return st;
  float *d = static_cast<float *>(scratch);
  launch_resolve_rgb(ctx->stream, film->accum, film->n_owned, scale, d);
  ctx->counters.kernel_launches += 1;
  cudaError_t e = cudaMemcpyAsync(host_rgb, d, (size_t)film->n_owned * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess)
    return rt_cuda_fail(e, "rt_film_read_rgb");
  return RT_OK;
}

int rt_film_resolve_rgb8_device(rt_film *film, double scale, void *device_rgb8) {
  if (!film || !device_rgb8)
    return invalid("rt_film_resolve_rgb8_device: null argument");
  RT_CUDA(cudaSetDevice(film->ctx->device));
  if (int flushed = rt_film_flush(film))
    return flushed;
  // compact order out (rank 0 of 1): the staged 16-byte-store kernel of rt_frame.cu
  launch_present_rgb8(film->ctx, film->ctx->stream, film->accum, film->n_owned, film->map.width, film->map.tile_rows, 0, 1,
                      scale, (uint8_t *)device_rgb8, nullptr, nullptr, 0, nullptr, nullptr);
  film->ctx->counters.kernel_launches += 1;
  RT_CUDA(cudaGetLastError());
  return RT_OK;
}

int rt_film_resolve_rgb8(rt_film *film, double scale, uint8_t *host_rgb8) {
  if (!film || !host_rgb8)
    return invalid("rt_film_resolve_rgb8: null argument");
  if (film->n_owned == 0)
    return RT_OK;
  rt_context *ctx = film->ctx;
  RT_CUDA(cudaSetDevice(ctx->device));
  void *scratch = nullptr;
  int st = ctx_scratch(ctx, (size_t)film->n_owned * 3, &scratch);
  if (st != RT_OK)
    return st;
  uint8_t *d = static_cast<uint8_t *>(scratch);
  st = rt_film_resolve_rgb8_device(film, scale, d);
  cudaError_t e = cudaSuccess;
  if (st == RT_OK) {
    e = cudaMemcpyAsync(host_rgb8, d, (size_t)film->n_owned * 3, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(ctx->stream);
  }
  if (st != RT_OK)
    return st;
  if (e != cudaSuccess)
    return rt_cuda_fail(e, "rt_film_resolve_rgb8");
  return RT_OK;
}

static int scatter_gathered(rt_context *ctx, int width, int height, int n_ranks, int tile_rows, const void *device_gathered,
                            void *device_full_image, int bytes_per_pixel) {
  if (!ctx || !device_gathered || !device_full_image || width < 1 || height < 1 || n_ranks < 1 || n_ranks > 64 ||
      tile_rows < 1)
    return invalid("rt_film_scatter_gathered: bad argument (1..64 ranks)");
  RT_CUDA(cudaSetDevice(ctx->device));
  launch_scatter_gathered(ctx->stream, width, height, n_ranks, tile_rows, device_gathered, device_full_image,
                          bytes_per_pixel);
  ctx->counters.kernel_launches += 1;
  RT_CUDA(cudaGetLastError());
  return RT_OK;
}

int rt_film_scatter_gathered(rt_context *ctx, int width, int height, int n_ranks, int tile_rows,
                             const void *device_gathered, void *device_full_image) {
  return scatter_gathered(ctx, width, height, n_ranks, tile_rows, device_gathered, device_full_image, 16);
}

int rt_film_scatter_gathered_rgb8(rt_context *ctx, int width, int height, int n_ranks, int tile_rows,
                                  const void *device_gathered, void *device_full_image) {
  return scatter_gathered(ctx, width, height, n_ranks, tile_rows, device_gathered, device_full_image, 3);
}

// Single-process multi-GPU gather: every rank's tiles are copied over NVLink (peer copies) into one rank-major
// buffer on rank 0's device and scattered into the full image there.  as_rgb8: every rank tone-maps its own
// tiles first (to_byte, FP64 - the bytes are identical to a single-GPU resolve) and 3 bytes per pixel travel;
// otherwise the float4 sums travel and rank 0 scales them.
static int gather_p2p(rt_film **films, int n_ranks, double scale, float *host_rgb, uint8_t *host_rgb8) {
  const bool as_rgb8 = host_rgb8 != nullptr;
  if (!films || n_ranks < 1 || n_ranks > 64 || (!host_rgb && !host_rgb8))
    return invalid("rt_film_gather_p2p: bad argument (1..64 ranks)");
  for (int r = 0; r < n_ranks; r++) {
    if (!films[r] || films[r]->map.rank != r || films[r]->map.n_ranks != n_ranks ||
        films[r]->map.width != films[0]->map.width || films[r]->map.height != films[0]->map.height ||
        films[r]->map.tile_rows != films[0]->map.tile_rows)
      return invalid("rt_film_gather_p2p: films[r] must be rank r of n_ranks with one geometry");
  }
  rt_context *ctx0 = films[0]->ctx;
  const int W = films[0]->map.width, H = films[0]->map.height;
  const int64_t total = (int64_t)W * H;
  const size_t px = as_rgb8 ? 3 : sizeof(float4);
  std::vector<void *> staged(n_ranks, nullptr); // per-rank RGB8 tiles (as_rgb8 only)
  unsigned char *gathered = nullptr, *full = nullptr;
  int st = RT_OK;
  cudaError_t e = cudaSuccess;
  auto cleanup = [&]() {
    for (int r = 0; r < n_ranks; r++)
      if (staged[r]) {
        cudaSetDevice(films[r]->ctx->device);
        cudaFree(staged[r]);
      }
    cudaSetDevice(ctx0->device);
    cudaFree(gathered);
    cudaFree(full);
  };
  for (int r = 0; r < n_ranks && st == RT_OK; r++) {
    RT_CUDA(cudaSetDevice(films[r]->ctx->device));
    if ((st = rt_film_flush(films[r])) != RT_OK)
      break;
    if (as_rgb8 && films[r]->n_owned > 0) {
      e = cudaMalloc(&staged[r], (size_t)films[r]->n_owned * 3);
      if (e != cudaSuccess)
        st = rt_cuda_fail(e, "cudaMalloc (rgb8 tiles)");
      else
        st = rt_film_resolve_rgb8_device(films[r], scale, staged[r]);
    }
    if (st == RT_OK && (e = cudaStreamSynchronize(films[r]->ctx->stream)) != cudaSuccess)
      st = rt_cuda_fail(e, "cudaStreamSynchronize");
    if (r > 0) { // direct NVLink copies into rank 0 (an error here only means it was enabled before)
      cudaSetDevice(ctx0->device);
      if (cudaDeviceEnablePeerAccess(films[r]->ctx->device, 0) != cudaSuccess)
        cudaGetLastError();
    }
  }
  if (st != RT_OK) {
    cleanup();
    return st;
  }
  cudaSetDevice(ctx0->device);
  if ((e = cudaMalloc((void **)&gathered, (size_t)total * px)) != cudaSuccess ||
      (e = cudaMalloc((void **)&full, (size_t)total * px)) != cudaSuccess) {
    cleanup();
    return rt_cuda_fail(e, "cudaMalloc (gather)");
  }
  int64_t offset = 0;
  for (int r = 0; r < n_ranks && e == cudaSuccess; r++) {
    size_t bytes = (size_t)films[r]->n_owned * px;
    const void *src = as_rgb8 ? staged[r] : (const void *)films[r]->accum;
    if (bytes)
      e = cudaMemcpyPeerAsync(gathered + (size_t)offset * px, ctx0->device, src, films[r]->ctx->device, bytes, ctx0->stream);
    offset += films[r]->n_owned;
  }
  if (e != cudaSuccess) {
    st = rt_cuda_fail(e, "cudaMemcpyPeerAsync");
  } else {
    launch_scatter_gathered(ctx0->stream, W, H, n_ranks, films[0]->map.tile_rows, gathered, full, (int)px);
    ctx0->counters.kernel_launches += 1;
    if (as_rgb8) {
      e = cudaMemcpyAsync(host_rgb8, full, (size_t)total * 3, cudaMemcpyDeviceToHost, ctx0->stream);
      if (e == cudaSuccess)
        e = cudaStreamSynchronize(ctx0->stream);
      if (e != cudaSuccess)
        st = rt_cuda_fail(e, "rt_film_gather_p2p_rgb8");
    } else {
      rt_film whole;
      whole.ctx = ctx0;
      whole.accum = (float4 *)full;
      whole.n_owned = total;
      st = rt_film_read_rgb(&whole, scale, host_rgb);
    }
  }
  cudaStreamSynchronize(ctx0->stream);
  cleanup();
  return st;
}

int rt_film_gather_p2p(rt_film **films, int n_ranks, double scale, float *host_rgb) {
  if (!host_rgb)
    return invalid("rt_film_gather_p2p: null output");
  return gather_p2p(films, n_ranks, scale, host_rgb, nullptr);
}

int rt_film_gather_p2p_rgb8(rt_film **films, int n_ranks, double scale, uint8_t *host_rgb8) {
  if (!host_rgb8)
    return invalid("rt_film_gather_p2p_rgb8: null output");
  return gather_p2p(films, n_ranks, scale, nullptr, host_rgb8);
}

int rt_get_counters(rt_context *ctx, rt_counters *out) {
  if (!ctx || !out)
    return invalid("rt_get_counters: null argument");
  RT_CUDA(cudaSetDevice(ctx->device));
  unsigned long long stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ctx->wave.stats) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    RT_CUDA(cudaMemcpy(stats, ctx->wave.stats, sizeof stats, cudaMemcpyDeviceToHost));
  }
  *out = ctx->counters;
  out->segments = stats[0] + stats[3]; // wavefront extend launches + tail kernel
  out->tail_segments = stats[3];
  out->nodes_visited = stats[1]; // counted by the instrumented kernels only (rt_context_set_stats)
  out->prim_tests = stats[2];
  if (std::getenv("RT_DEBUG_STATS")) // development aid: the longest single traversal the instrumented kernels saw
    std::fprintf(stderr, "[rt stats] longest traversal: %llu node visits; rays above 512 node visits: %llu\n", stats[4], stats[5]);
  return RT_OK;
}

int rt_context_set_graph(rt_context *ctx, int enable) {
  if (!ctx)
    return invalid("null context");
  ctx->use_graph = enable != 0;
  return RT_OK;
}

int rt_context_set_stats(rt_context *ctx, int enable) {
  if (!ctx)
    return invalid("null context");
  ctx->stats = enable != 0;
  return RT_OK;
}

int rt_get_queue_lengths(rt_context *ctx, uint32_t *out, int n) {
  if (!ctx || !out || n < 0)
    return invalid("rt_get_queue_lengths: bad argument");
  RT_CUDA(cudaSetDevice(ctx->device));
  for (int k = 0; k < n; k++)
    out[k] = 0;
  // the first half of `counts` holds the queue lengths, the second the fetch cursors
  size_t have = std::min<size_t>((size_t)n, ctx->wave.last_counts);
  if (have && ctx->wave.counts) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> part(have);
    for (size_t h = 0; h < ctx->wave.last_splits; h++) { // a split pass: the sub-passes' queues add up
      RT_CUDA(cudaMemcpy(part.data(), ctx->wave.counts + h * ctx->wave.last_stride, have * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      for (size_t k = 0; k < have; k++)
        out[k] += part[k];
    }
  }
  return RT_OK;
}

int rt_context_set_audit(rt_context *ctx, int enable) {
  if (!ctx)
    return invalid("null context");
  RT_CUDA(cudaSetDevice(ctx->device));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->audit = enable != 0;
  if (ctx->wave.audit_stats)
    RT_CUDA(cudaMemsetAsync(ctx->wave.audit_stats, 0, RT_AUDIT_WORDS * sizeof(unsigned long long), ctx->stream));
  return RT_OK;
}

int rt_get_audit(rt_context *ctx, rt_audit *out) {
  if (!ctx || !out)
    return invalid("rt_get_audit: null argument");
  RT_CUDA(cudaSetDevice(ctx->device));
  unsigned long long t[RT_AUDIT_WORDS] = {};
  if (ctx->wave.audit_stats) {
    RT_CUDA(cudaStreamSynchronize(ctx->stream));
    RT_CUDA(cudaMemcpy(t, ctx->wave.audit_stats, sizeof t, cudaMemcpyDeviceToHost));
  }
  std::memset(out, 0, sizeof *out);
  out->segments = t[RT_AUDIT_SEGMENTS];
  out->prim_mismatch = t[RT_AUDIT_MISMATCH];
  out->primary_segments = t[RT_AUDIT_PRIMARY];
  out->primary_mismatch = t[RT_AUDIT_PRIMARY_MISMATCH];
  out->hit_miss_flips = t[RT_AUDIT_HIT_MISS];
  out->t_rel_above_1e4 = t[RT_AUDIT_T_ABOVE_1E4];
  uint32_t bits = (uint32_t)t[RT_AUDIT_MAX_REL_T];
  float f;
  std::memcpy(&f, &bits, 4);
  out->max_rel_t_error = f;
  out->rechecked = 0;
  return RT_OK;
}

int rt_get_audit_samples(rt_context *ctx, rt_audit_sample *out, int max_samples) {
  if (!ctx || (max_samples > 0 && !out) || max_samples < 0) {
    invalid("rt_get_audit_samples: bad argument");
    return -1;
  }
  if (!ctx->wave.audit_stats || cudaSetDevice(ctx->device) != cudaSuccess)
    return 0;
  unsigned long long n = 0;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess ||
      cudaMemcpy(&n, ctx->wave.audit_stats + RT_AUDIT_SAMPLES, sizeof n, cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  int have = (int)std::min<unsigned long long>(n, std::min<unsigned long long>(RT_AUDIT_MAX_SAMPLES, (unsigned long long)max_samples));
  if (have && cudaMemcpy(out, ctx->wave.audit_samples, (size_t)have * sizeof(rt_audit_sample), cudaMemcpyDeviceToHost) != cudaSuccess)
    return -1;
  return have;
}

int rt_context_set_stage_timing(rt_context *ctx, int enable) {
  if (!ctx)
    return invalid("null context");
  RT_CUDA(cudaSetDevice(ctx->device));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->timer.enabled = enable != 0;
  ctx->timer.spans.clear();
  ctx->timer.next = 0;
  return RT_OK;
}

int rt_get_stage_times(rt_context *ctx, double ms[RT_STAGE_COUNT], uint64_t launches[RT_STAGE_COUNT]) {
  if (!ctx || !ms || !launches)
    return invalid("rt_get_stage_times: null argument");
  RT_CUDA(cudaSetDevice(ctx->device));
  RT_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int k = 0; k < RT_STAGE_COUNT; k++) {
    ms[k] = 0.0;
    launches[k] = 0;
  }
  for (const auto &span : ctx->timer.spans) {
    float t = 0.f;
    RT_CUDA(cudaEventElapsedTime(&t, span.second.first, span.second.second));
    ms[span.first] += t;
    launches[span.first] += 1;
  }
  ctx->timer.spans.clear();
  ctx->timer.next = 0;
  return RT_OK;
}

int rt_reset_counters(rt_context *ctx) {
  if (!ctx)
    return invalid("null context");
  RT_CUDA(cudaSetDevice(ctx->device));
  ctx->counters = rt_counters{};
  if (ctx->wave.stats)
    RT_CUDA(cudaMemsetAsync(ctx->wave.stats, 0, 8 * sizeof(unsigned long long), ctx->stream));
  return RT_OK;
}

} // extern "C"
