// rt_scene.cu — scene upload: flat description -> device SoA records + GPU-built BVH4.
//
// Replaces initialize_cuda_scene and the per-object converters (scene/CudaSceneInitialization.cuh:249-299,
// core/HittableConverter.cuh:50-111): instead of mirroring the object graph node by node with one
// cudaMemcpy each, instance chains are baked into world-space primitives on the host in FP64, every
// array is uploaded once, and the hierarchy is built on the device (Morton sort + Karras + refit +
// 4-wide collapse, rt_bvh.h).
#include "rt_internal.h"

#include <cstdio>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "rt_flatten.h"
#include "rt_sah.h"

#include <chrono>
#include <mutex>
#include <string>
#include <thread>

using namespace rtflat;

namespace {

template <class T, class Vec> int upload(T **dst, const Vec &src, size_t min_count = 1) {
  size_t n = std::max(src.size(), min_count);
  RT_CUDA(cudaMalloc((void **)dst, n * sizeof(T)));
  if (src.size() < n)
    RT_CUDA(cudaMemset(*dst, 0, n * sizeof(T)));
  if (!src.empty())
    RT_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return RT_OK;
}

struct Scratch { // frees build scratch on every exit path
  std::vector<void *> ptrs;
  ~Scratch() {
    for (void *p : ptrs)
      cudaFree(p);
  }
  template <class T> cudaError_t alloc(T **p, size_t count) {
    cudaError_t e = cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T));
    if (e == cudaSuccess)
      ptrs.push_back(*p);
    return e;
  }
};

} // namespace

static int scene_build_once(rt_context *ctx, const rt_scene_desc *desc, rt_scene *sc) {
  // RT_BUILD_TIMING=1: wall-clock phases of the call on stderr (where scene creation time goes)
  static const bool timing = std::getenv("RT_BUILD_TIMING") != nullptr;
  auto wall0 = std::chrono::steady_clock::now();
  auto phase = [&](const char *what) {
    if (timing) {
      cudaStreamSynchronize(ctx->stream);
      auto now = std::chrono::steady_clock::now();
      std::fprintf(stderr, "[rt_scene_build] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(now - wall0).count());
      wall0 = now;
    }
  };
  Flat f;
  int st = flatten(desc, f);
  if (st != RT_OK)
    return st;
  phase("flatten (host, FP64 bake)");
  const int n = (int)f.boxes.size();
  sc->ctx = ctx;
  sc->n_leaf = n;
  cudaStream_t s = ctx->stream;

  // Two helper threads work beside the build (joined on every way out of this function):
  //  * everything that does not depend on leaf order is uploaded while the host prepares the leaf boxes (a million
  //    materials are 48 MB of pageable memory: a synchronous copy of several milliseconds);
  //  * the host-side copies of the description that rt_scene_update_* needs later are plain memory traffic.
  struct Helper {
    std::thread t;
    ~Helper() {
      if (t.joinable())
        t.join();
    }
  };
  int side_status = RT_OK;
  std::string side_error;
  Helper side_upload, side_copies;
  side_upload.t = std::thread([&]() {
    int s2 = RT_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess)
      s2 = RT_ERR_CUDA;
    else if ((s2 = upload(&sc->bprims, f.bprims)) || (s2 = upload(&sc->mats, f.mats)) || (s2 = upload(&sc->lights, f.lights)) ||
             (s2 = upload(&sc->perlin_grad, f.perlin_grad)) || (s2 = upload(&sc->perlin_perm, f.perlin_perm)) ||
             (s2 = upload(&sc->texels, f.texels)) || (s2 = upload(&sc->ex_bprims, f.ex_bprims)) ||
             (s2 = upload(&sc->ex_ops, f.ops)) || (s2 = upload(&sc->ex_chain_first, f.chain_first)) ||
             (s2 = upload(&sc->ex_chain_count, f.chain_count))) {
    }
    if (s2 != RT_OK)
      side_error = rt_last_error(); // the message is thread-local: carry it over to the caller's thread
    side_status = s2;
  });
  side_copies.t = std::thread([&]() {
    sc->h_xforms.assign(desc->xforms, desc->xforms + desc->n_xforms);
    sc->h_xform_ops.assign(desc->xform_ops, desc->xform_ops + desc->n_xform_ops);
    sc->h_spheres.resize(desc->n_spheres);
    parallel_for((size_t)desc->n_spheres, [&](size_t a, size_t b) { std::copy(desc->spheres + a, desc->spheres + b, sc->h_spheres.begin() + a); });
    sc->h_quads.resize(desc->n_quads);
    parallel_for((size_t)desc->n_quads, [&](size_t a, size_t b) { std::copy(desc->quads + a, desc->quads + b, sc->h_quads.begin() + a); });
  });

  BigVec<BuildBox> boxes;
  boxes.resize(n);
  BoxD all, centroids;
  {
    std::mutex merge;
    parallel_for((size_t)n, [&](size_t a, size_t b) {
      BoxD my_all, my_centroids;
      for (size_t i = a; i < b; i++) {
        boxes[i] = to_build_box(f.boxes[i]);
        my_all.grow(f.boxes[i]);
        my_centroids.grow(D3{0.5 * (boxes[i].lo[0] + boxes[i].hi[0]), 0.5 * (boxes[i].lo[1] + boxes[i].hi[1]),
                             0.5 * (boxes[i].lo[2] + boxes[i].hi[2])});
      }
      std::lock_guard<std::mutex> lock(merge);
      all.grow(my_all);
      centroids.grow(my_centroids);
    });
  }

  phase("leaf boxes (host)");
  const int n_wide_cap = std::max(n, 1);
  RT_CUDA(cudaMalloc((void **)&sc->nodes, (size_t)n_wide_cap * RT_NODE_F4 * sizeof(float4)));
  RT_CUDA(cudaMalloc((void **)&sc->prims, (size_t)std::max(n, 1) * RT_PRIM_F4 * sizeof(float4)));
  RT_CUDA(cudaMalloc((void **)&sc->leaf_object, (size_t)std::max(n, 1) * sizeof(int)));
  RT_CUDA(cudaMalloc((void **)&sc->leaf_id, (size_t)std::max(n, 1) * sizeof(int)));

  cudaEvent_t ev0, ev1;
  RT_CUDA(cudaEventCreate(&ev0));
  RT_CUDA(cudaEventCreate(&ev1));
  int n_wide = 1;
  BigVec<uint32_t> order; // written as a whole by the download of the tree's leaf order
  order.resize(n);

  if (n <= 1) {
    // degenerate trees: one wide node with zero or one leaf child
    std::vector<float4> node(RT_NODE_F4);
    const float inf = std::numeric_limits<float>::infinity();
    for (int a = 0; a < 3; a++) {
      node[2 * a] = make_float4(n ? boxes[0].lo[a] : inf, inf, inf, inf);
      node[2 * a + 1] = make_float4(n ? boxes[0].hi[a] : -inf, -inf, -inf, -inf);
    }
    node[6] = make_float4(ibits(n ? ~0 : RT_EMPTY), ibits(RT_EMPTY), ibits(RT_EMPTY), ibits(RT_EMPTY));
    node[7] = make_float4(ibits(-1), 0.f, 0.f, 0.f); // the root has no parent slot
    RT_CUDA(cudaMemcpy(sc->nodes, node.data(), sizeof(float4) * RT_NODE_F4, cudaMemcpyHostToDevice));
    if (n) {
      RT_CUDA(cudaMemcpy(sc->prims, f.prims.data(), sizeof(float4) * RT_PRIM_F4, cudaMemcpyHostToDevice));
      order[0] = 0;
    }
    sc->info.build_ms = 0.0;
    sc->info.depth = 1;
  } else {
    Scratch scratch;
    BuildBox *d_boxes = nullptr, *d_sorted_boxes = nullptr;
    float4 *d_prims_in = nullptr;
    uint64_t *d_codes = nullptr, *d_codes_sorted = nullptr;
    uint32_t *d_index = nullptr, *d_index_sorted = nullptr;
    float *d_bounds = nullptr;
    BinTree t{};
    CollapseItem *d_items[2] = {nullptr, nullptr};
    int *d_counters = nullptr; // [0] next queue length  [1] wide node count
    RT_CUDA(scratch.alloc(&d_boxes, n));
    RT_CUDA(scratch.alloc(&d_sorted_boxes, n));
    RT_CUDA(scratch.alloc(&d_prims_in, (size_t)n * RT_PRIM_F4));
    RT_CUDA(scratch.alloc(&d_codes, n));
    RT_CUDA(scratch.alloc(&d_codes_sorted, n));
    RT_CUDA(scratch.alloc(&d_index, n));
    RT_CUDA(scratch.alloc(&d_index_sorted, n));
    RT_CUDA(scratch.alloc(&d_bounds, 6));
    RT_CUDA(scratch.alloc(&t.left, n - 1));
    RT_CUDA(scratch.alloc(&t.right, n - 1));
    RT_CUDA(scratch.alloc(&t.parent, 2 * n - 1));
    RT_CUDA(scratch.alloc(&t.box, n - 1));
    RT_CUDA(scratch.alloc(&t.visits, n - 1));
    RT_CUDA(scratch.alloc(&d_items[0], n));
    RT_CUDA(scratch.alloc(&d_items[1], n));
    RT_CUDA(scratch.alloc(&d_counters, 2));
    t.n = n;

    float bounds[6];
    for (int a = 0; a < 3; a++) {
      bounds[a] = (float)centroids.lo[a];
      double ext = centroids.hi[a] - centroids.lo[a];
      bounds[3 + a] = ext > 0 ? (float)(1.0 / ext) : 0.f;
    }
    RT_CUDA(cudaMemcpyAsync(d_boxes, boxes.data(), sizeof(BuildBox) * n, cudaMemcpyHostToDevice, s));
    RT_CUDA(cudaMemcpyAsync(d_prims_in, f.prims.data(), sizeof(float4) * RT_PRIM_F4 * n, cudaMemcpyHostToDevice, s));
    RT_CUDA(cudaMemcpyAsync(d_bounds, bounds, sizeof bounds, cudaMemcpyHostToDevice, s));
    RT_CUDA(cudaMemsetAsync(t.visits, 0, sizeof(unsigned int) * (n - 1), s));

    phase("uploads");
    // Which binary tree: the host SAH tree (rt_sah.h) from 64 to 65,536 primitives, a device tree otherwise (the
    // Karras radix tree; PLOC with RT_BVH=ploc).  RT_BVH = sah / lbvh / ploc forces one builder; RT_BVH=best
    // builds the host SAH tree AND the PLOC tree and keeps the smaller surface-area sum (an experiment aid: the
    // area sum picks the tree with fewer node visits, which was not the faster one on the final scene).
    const char *bvh_env = std::getenv("RT_BVH");
    const bool best_of = bvh_env && !std::strcmp(bvh_env, "best") && n >= RT_SAH_MIN_PRIMS && n <= RT_SAH_MAX_PRIMS;
    const bool host_sah = rtsah::use_sah(n);
    const bool device_tree = !host_sah || best_of;
    const bool ploc = rtsah::use_ploc(n) || best_of;
    double host_build_ms = 0.0, host_area = 0.0;
    if (host_sah) {
      auto t0 = std::chrono::steady_clock::now();
      rtsah::HostTree ht;
      rtsah::build(boxes.data(), n, ht);
      for (const BuildBox &b : ht.box)
        host_area += (double)box_area(b);
      host_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      // uploaded in the layout the device stages produce (children, boxes, primitive order)
      RT_CUDA(cudaMemcpyAsync(d_index_sorted, ht.order.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
      RT_CUDA(cudaMemcpyAsync(t.left, ht.left.data(), sizeof(int) * (n - 1), cudaMemcpyHostToDevice, s));
      RT_CUDA(cudaMemcpyAsync(t.right, ht.right.data(), sizeof(int) * (n - 1), cudaMemcpyHostToDevice, s));
      RT_CUDA(cudaMemcpyAsync(t.box, ht.box.data(), sizeof(BuildBox) * (n - 1), cudaMemcpyHostToDevice, s));
      RT_CUDA(cudaStreamSynchronize(s)); // ht goes out of scope
    }
    RT_CUDA(cudaEventRecord(ev0, s));
    sc->info.builder = host_sah ? RT_BUILDER_SAH : (ploc ? RT_BUILDER_PLOC : RT_BUILDER_LBVH);
    if (device_tree) {
      // the device candidate: Morton order, then PLOC rounds (or the Karras hierarchy + refit)
      BinTree dt = t;
      uint32_t *d_order = d_index_sorted;
      if (host_sah) { // a second set of tree arrays beside the uploaded host tree
        RT_CUDA(scratch.alloc(&dt.left, n - 1));
        RT_CUDA(scratch.alloc(&dt.right, n - 1));
        RT_CUDA(scratch.alloc(&dt.parent, 2 * n - 1));
        RT_CUDA(scratch.alloc(&dt.box, n - 1));
        RT_CUDA(scratch.alloc(&d_order, n));
      }
      launch_morton(s, d_boxes, n, d_bounds, d_bounds + 3, d_codes, d_index);
      if ((st = sort_pairs(s, d_codes, d_codes_sorted, d_index, d_order, n)))
        return st;
      launch_gather_boxes(s, d_boxes, d_order, d_sorted_boxes, n);
      if (ploc) {
        PlocCluster *clusters[2] = {nullptr, nullptr};
        int *d_nearest = nullptr;
        unsigned long long *d_packed = nullptr;
        RT_CUDA(scratch.alloc(&clusters[0], n));
        RT_CUDA(scratch.alloc(&clusters[1], n));
        RT_CUDA(scratch.alloc(&d_nearest, n));
        RT_CUDA(scratch.alloc(&d_packed, n));
        int rounds = 0;
        if ((st = ploc_build(s, d_sorted_boxes, n, dt, clusters, d_nearest, d_packed, &rounds)))
          return st;
      } else {
        launch_hierarchy(s, d_codes_sorted, dt);
        launch_refit(s, dt, d_sorted_boxes);
      }
      bool keep_device = true;
      if (host_sah) {
        double *d_sum = nullptr, device_area = 0.0;
        RT_CUDA(scratch.alloc(&d_sum, 1));
        if ((st = tree_area(s, dt.box, n - 1, d_sum, &device_area)))
          return st;
        keep_device = device_area < host_area;
      }
      if (keep_device) {
        t = dt;
        d_index_sorted = d_order;
        sc->info.builder = ploc ? RT_BUILDER_PLOC : RT_BUILDER_LBVH;
      }
    }
    phase("binary tree");
    // leaf boxes and records in the order of the tree that was kept
    launch_gather_boxes(s, d_boxes, d_index_sorted, d_sorted_boxes, n);
    launch_gather_records(s, d_prims_in, d_index_sorted, sc->prims, n, RT_PRIM_F4 * (int)sizeof(float4));

    // collapse, level by level; the root binary node 0 becomes wide node 0
    CollapseItem root{0, 0, -1};
    int counters[2] = {0, 1};
    RT_CUDA(cudaMemcpyAsync(d_items[0], &root, sizeof root, cudaMemcpyHostToDevice, s));
    RT_CUDA(cudaMemcpyAsync(d_counters, counters, sizeof counters, cudaMemcpyHostToDevice, s));
    int n_items = 1, cur = 0, levels = 0;
    while (n_items > 0) {
      levels++; // one collapse launch per level of the 4-wide tree
      RT_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(int), s)); // next level's queue length
      launch_collapse(s, t, d_sorted_boxes, sc->nodes, d_items[cur], n_items, d_items[cur ^ 1], d_counters,
                      d_counters + 1);
      RT_CUDA(cudaMemcpyAsync(counters, d_counters, sizeof counters, cudaMemcpyDeviceToHost, s));
      RT_CUDA(cudaStreamSynchronize(s));
      n_items = counters[0];
      cur ^= 1;
    }
    n_wide = counters[1];
    sc->info.depth = levels;
    RT_CUDA(cudaEventRecord(ev1, s));
    RT_CUDA(cudaEventSynchronize(ev1));
    float ms = 0.f;
    RT_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    sc->info.build_ms = ms + host_build_ms;
    RT_CUDA(cudaMemcpy(order.data(), d_index_sorted, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    RT_CUDA(cudaGetLastError());
    phase("gathers + collapse");
  }
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);

  // per-leaf object / id tables in leaf order
  BigVec<int> leaf_object, leaf_id;
  leaf_object.resize(std::max(n, 1));
  leaf_id.resize(std::max(n, 1));
  leaf_object[0] = leaf_id[0] = -1; // the placeholder entry of an empty scene
  parallel_for((size_t)n, [&](size_t a, size_t b) {
    for (size_t j = a; j < b; j++) {
      leaf_object[j] = f.ex_prims[order[j]].object;
      leaf_id[j] = f.ex_prims[order[j]].id;
    }
  });
  RT_CUDA(cudaMemcpy(sc->leaf_object, leaf_object.data(), sizeof(int) * leaf_object.size(), cudaMemcpyHostToDevice));
  RT_CUDA(cudaMemcpy(sc->leaf_id, leaf_id.data(), sizeof(int) * leaf_id.size(), cudaMemcpyHostToDevice));

  // kept on the host for rt_scene_update_spheres (the description copies: side_copies above)
  side_upload.t.join();
  if (side_status != RT_OK) {
    rt_set_error(side_error);
    return side_status;
  }
  sc->h_mats = std::move(f.mats); // uploaded; not needed by this function any more
  // The FP64 parity records (160 B per primitive) are only read by rt_trace_rays(RT_TRACE_EXACT_F64), the parity
  // audit and primitive updates: they stay on the host until one of those asks (rt_scene_ensure_exact).
  sc->h_ex_prims = std::move(f.ex_prims); // description order; h_order[j] = record of leaf j
  {
    BigVec<int> leaf_of_record; // a permutation: every entry below n is written
    leaf_of_record.resize(std::max(n, 1));
    parallel_for((size_t)n, [&](size_t a, size_t b) {
      for (size_t j = a; j < b; j++)
        leaf_of_record[order[j]] = (int)j;
    });
    sc->sphere_leaf.assign(desc->n_spheres, -1);
    sc->quad_leaf.assign(desc->n_quads, -1);
    int record = 0; // surface spheres are the first records, then the surface quads, in description order (rt_flatten.h)
    for (int i = 0; i < desc->n_spheres; i++)
      if (!(desc->spheres[i].flags & RT_PRIM_BOUNDARY))
        sc->sphere_leaf[i] = leaf_of_record[record++];
    for (int i = 0; i < desc->n_quads; i++)
      if (!(desc->quads[i].flags & RT_PRIM_BOUNDARY))
        sc->quad_leaf[i] = leaf_of_record[record++];
  }
  sc->h_order = std::move(order);
  side_copies.t.join();

  sc->d.nodes = sc->nodes;
  sc->d.prims = sc->prims;
  sc->d.bprims = sc->bprims;
  sc->d.mats = sc->mats;
  sc->d.lights = sc->lights;
  sc->d.perlin_grad = sc->perlin_grad;
  sc->d.perlin_perm = sc->perlin_perm;
  sc->d.texels = sc->texels;
  sc->d.n_prims = n;
  sc->d.n_lights = desc->n_lights;
  sc->d.n_media = desc->n_media;
  sc->d.bg[0] = sc->d.bg[1] = sc->d.bg[2] = 0.f; // set per render from the camera
  sc->ex.nodes = sc->nodes;
  sc->ex.prims = nullptr; // rt_scene_ensure_exact
  sc->ex.bprims = sc->ex_bprims;
  sc->ex.ops = sc->ex_ops;
  sc->ex.chain_first = sc->ex_chain_first;
  sc->ex.chain_count = sc->ex_chain_count;

  phase("leaf tables + host copies");
  sc->info.n_prims = n;
  sc->info.n_nodes = n_wide;
  sc->info.node_bytes = (int64_t)n_wide * RT_NODE_F4 * (int64_t)sizeof(float4);
  sc->info.prim_bytes = (int64_t)n * RT_PRIM_F4 * (int64_t)sizeof(float4);
  for (int a = 0; a < 3; a++) {
    sc->info.bounds_min[a] = n ? all.lo[a] : 0.0;
    sc->info.bounds_max[a] = n ? all.hi[a] : 0.0;
  }
  return RT_OK;
}

// A traversal descends at most `depth` levels of the 4-wide tree and leaves at most three siblings on its stack per
// level, so 3 x depth <= RT_STACK entries are always enough.  Host SAH trees and radix trees over spread-out scenes are
// far below that (10^6 spheres: 12 levels); a radix tree over Morton codes that share long prefixes (geometry
// clustered at many scales) can exceed it, and a dropped push would silently lose geometry - such a scene is rebuilt
// with the host SAH builder, and refused if even that tree is too deep.
int rt_scene_build(rt_context *ctx, const rt_scene_desc *desc, rt_scene *sc) {
  int st = scene_build_once(ctx, desc, sc);
  if (st != RT_OK || 3 * sc->info.depth <= RT_STACK)
    return st;
  const int first_depth = sc->info.depth;
  if (sc->info.builder != RT_BUILDER_SAH) {
    cudaStreamSynchronize(ctx->stream);
    rt_scene_release(sc);
    *sc = rt_scene();
    rtsah::g_force_sah = true;
    st = scene_build_once(ctx, desc, sc);
    rtsah::g_force_sah = false;
    if (st != RT_OK || 3 * sc->info.depth <= RT_STACK)
      return st;
  }
  rt_set_error("scene hierarchy too deep for the traversal stack (" + std::to_string(first_depth) + " levels, then " +
               std::to_string(sc->info.depth) + " with the SAH builder; at most " + std::to_string(RT_STACK / 3) + ")");
  return RT_ERR_UNSUPPORTED;
}

// Uploads the FP64 parity records on first use and brings them into leaf order on the device.
int rt_scene_ensure_exact(rt_scene *sc) {
  if (sc->ex_prims)
    return RT_OK;
  const size_t n = sc->h_order.size();
  cudaStream_t s = sc->ctx->stream;
  RT_CUDA(cudaMalloc((void **)&sc->ex_prims, std::max<size_t>(n, 1) * sizeof(PrimExact)));
  if (n) {
    PrimExact *d_in = nullptr;
    uint32_t *d_order = nullptr;
    RT_CUDA(cudaMalloc((void **)&d_in, n * sizeof(PrimExact)));
    cudaError_t e = cudaMalloc((void **)&d_order, n * sizeof(uint32_t));
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_in, sc->h_ex_prims.data(), n * sizeof(PrimExact), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d_order, sc->h_order.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
      static_assert(sizeof(PrimExact) % 16 == 0, "PrimExact must be a multiple of 16 bytes");
      launch_gather_records(s, d_in, d_order, sc->ex_prims, (int)n, (int)sizeof(PrimExact));
      e = cudaStreamSynchronize(s);
    }
    cudaFree(d_in);
    cudaFree(d_order);
    if (e != cudaSuccess)
      return rt_cuda_fail(e, "rt_scene_ensure_exact");
  }
  sc->ex.prims = sc->ex_prims;
  BigVec<PrimExact>().swap(sc->h_ex_prims); // the device copy is the master from here on
  BigVec<uint32_t>().swap(sc->h_order);
  return RT_OK;
}

// rt_scene_update_spheres / rt_scene_update_quads: re-bakes the given primitives (instance chains, material copy,
// FP64 parity record, box), scatters them to their leaves and refits the BVH4 bottom-up.  The tree keeps its topology.
static int update_primitives(rt_scene *sc, int first, int count, const rt_sphere *spheres, const rt_quad *quads) {
  const bool is_sphere = spheres != nullptr;
  const char *what = is_sphere ? "rt_scene_update_spheres" : "rt_scene_update_quads";
  const int n_have = is_sphere ? (int)sc->h_spheres.size() : (int)sc->h_quads.size();
  if (count == 0)
    return RT_OK;
  if ((!spheres && !quads) || first < 0 || count < 0 || first > n_have - count) {
    rt_set_error(std::string(what) + ": primitive range out of bounds");
    return RT_ERR_INVALID;
  }
  { // the update scatters FP64 parity records next to the render records
    int ready = rt_scene_ensure_exact(sc);
    if (ready != RT_OK)
      return ready;
  }
  rt_scene_desc d{};
  d.xforms = sc->h_xforms.data();
  d.n_xforms = (int)sc->h_xforms.size();
  d.xform_ops = sc->h_xform_ops.data();
  d.n_xform_ops = (int)sc->h_xform_ops.size();
  Baker bk{&d};
  const int n_materials = (int)(sc->h_mats.size() / RT_MAT_F4);
  const int n_spheres_total = (int)sc->h_spheres.size();
  std::vector<float4> records;
  std::vector<PrimExact> exact;
  std::vector<BuildBox> boxes;
  std::vector<int> leaves;
  for (int k = 0; k < count; k++) {
    const int i = first + k;
    const int leaf = is_sphere ? sc->sphere_leaf[i] : sc->quad_leaf[i];
    const int flags = is_sphere ? spheres[k].flags : quads[k].flags;
    const int xform = is_sphere ? spheres[k].xform : quads[k].xform;
    const int material = is_sphere ? spheres[k].material : quads[k].material;
    if (leaf < 0 || (flags & RT_PRIM_BOUNDARY)) {
      rt_set_error(std::string(what) + ": boundary primitives of media cannot be updated");
      return RT_ERR_UNSUPPORTED;
    }
    if (xform < -1 || xform >= d.n_xforms || material < 0 || material >= n_materials) {
      rt_set_error(std::string(what) + ": instance chain or material index out of range");
      return RT_ERR_INVALID;
    }
    BoxD box;
    if (is_sphere) {
      push_sphere(bk, spheres[k], i, material, records, exact, box);
      embed_sphere_material(&records[records.size() - RT_PRIM_F4], sc->h_mats);
    } else {
      push_quad(bk, quads[k], n_spheres_total + i, material, records, exact, box); // unified id: spheres first
    }
    boxes.push_back(to_build_box(box));
    leaves.push_back(leaf);
  }
  cudaStream_t st = sc->ctx->stream;
  const int n_nodes = (int)sc->info.n_nodes;
  if (!sc->leaf_up || !sc->arrivals) { // first update: where every leaf hangs, and the refit counters
    if (!sc->leaf_up)
      RT_CUDA(cudaMalloc((void **)&sc->leaf_up, sizeof(int) * std::max(sc->n_leaf, 1)));
    if (!sc->arrivals)
      RT_CUDA(cudaMalloc((void **)&sc->arrivals, sizeof(unsigned int) * std::max(n_nodes, 1)));
    launch_leaf_links(st, sc->nodes, n_nodes, sc->leaf_up);
  }
  Scratch scratch;
  float4 *d_records = nullptr;
  PrimExact *d_exact = nullptr;
  BuildBox *d_boxes = nullptr;
  int *d_leaves = nullptr;
  RT_CUDA(scratch.alloc(&d_records, records.size()));
  RT_CUDA(scratch.alloc(&d_exact, exact.size()));
  RT_CUDA(scratch.alloc(&d_boxes, boxes.size()));
  RT_CUDA(scratch.alloc(&d_leaves, leaves.size()));
  RT_CUDA(cudaMemcpyAsync(d_records, records.data(), sizeof(float4) * records.size(), cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaMemcpyAsync(d_exact, exact.data(), sizeof(PrimExact) * exact.size(), cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaMemcpyAsync(d_boxes, boxes.data(), sizeof(BuildBox) * boxes.size(), cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaMemcpyAsync(d_leaves, leaves.data(), sizeof(int) * leaves.size(), cudaMemcpyHostToDevice, st));
  RT_CUDA(cudaMemsetAsync(sc->arrivals, 0, sizeof(unsigned int) * std::max(n_nodes, 1), st));
  launch_update_leaves(st, d_records, d_exact, d_boxes, d_leaves, count, sc->leaf_up, sc->prims, sc->ex_prims, sc->nodes);
  launch_refit_wide(st, sc->nodes, sc->leaf_up, sc->arrivals, sc->n_leaf);
  RT_CUDA(cudaStreamSynchronize(st)); // the staging buffers go out of scope
  RT_CUDA(cudaGetLastError());
  if (is_sphere)
    std::copy(spheres, spheres + count, sc->h_spheres.begin() + first);
  else
    std::copy(quads, quads + count, sc->h_quads.begin() + first);
  return RT_OK;
}

int rt_scene_update_spheres_impl(rt_scene *sc, int first, int count, const rt_sphere *spheres) {
  if (!spheres && count != 0) {
    rt_set_error("rt_scene_update_spheres: null spheres");
    return RT_ERR_INVALID;
  }
  return update_primitives(sc, first, count, spheres, nullptr);
}
int rt_scene_update_quads_impl(rt_scene *sc, int first, int count, const rt_quad *quads) {
  if (!quads && count != 0) {
    rt_set_error("rt_scene_update_quads: null quads");
    return RT_ERR_INVALID;
  }
  return update_primitives(sc, first, count, nullptr, quads);
}

void rt_scene_release(rt_scene *sc) {
  cudaFree(sc->leaf_up);
  cudaFree(sc->arrivals);
  cudaFree(sc->nodes);
  cudaFree(sc->prims);
  cudaFree(sc->bprims);
  cudaFree(sc->mats);
  cudaFree(sc->lights);
  cudaFree(sc->perlin_grad);
  cudaFree(sc->perlin_perm);
  cudaFree(sc->texels);
  cudaFree(sc->ex_prims);
  cudaFree(sc->ex_bprims);
  cudaFree(sc->ex_ops);
  cudaFree(sc->ex_chain_first);
  cudaFree(sc->ex_chain_count);
  cudaFree(sc->leaf_object);
  cudaFree(sc->leaf_id);
}
