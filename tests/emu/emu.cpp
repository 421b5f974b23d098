// tests/emu/emu.cpp — TEST BUILD ONLY: the per-thread device functions of csrc/ (rt_device.h,
// rt_exact.h, rt_bvh.h, rt_flatten.h) compiled by the host compiler and driven by serial loops, so
// the flattening, the LBVH build logic, both traversals and the shading arithmetic can be checked
// against the oracle in a container without a GPU.  Nothing here is part of librt_b200.so, and no
// product code path can reach it; the GPU tests exercise the real kernels through the C ABI.
#include <cstdint>
static uint64_t g_stat_nodes = 0, g_stat_leaves = 0;
#define RT_STAT_NODE() (++g_stat_nodes)
#define RT_STAT_LEAF() (++g_stat_leaves)
#include "../../real-time-ray-tracing-engine_b200/csrc/rt_flatten.h"
#include "../../real-time-ray-tracing-engine_b200/csrc/rt_sah.h"

#include <algorithm>
#include <cstdio>
#include <numeric>
#include <string>
#include <vector>

static std::string g_err;
void rt_set_error(const std::string &msg) { g_err = msg; }

struct EmuScene {
  rtflat::Flat flat;
  std::vector<float4> nodes, prims;
  std::vector<PrimExact> ex_prims;
  std::vector<int> leaf_object, leaf_id;
  DScene d{};
  ExactScene ex{};
  int n_wide = 0;
  const rt_scene_desc *desc = nullptr; // caller-owned, alive as long as the scene (tests)
  std::vector<int> sphere_leaf, quad_leaf;
};

// The build stages of rt_scene.cu, one serial loop per kernel.
static void build_bvh(EmuScene &s) {
  using namespace rtflat;
  Flat &f = s.flat;
  const int n = (int)f.boxes.size();
  std::vector<BuildBox> boxes(n);
  BoxD centroids;
  for (int i = 0; i < n; i++) {
    boxes[i] = to_build_box(f.boxes[i]);
    centroids.grow(D3{0.5 * (boxes[i].lo[0] + boxes[i].hi[0]), 0.5 * (boxes[i].lo[1] + boxes[i].hi[1]),
                      0.5 * (boxes[i].lo[2] + boxes[i].hi[2])});
  }
  s.nodes.assign((size_t)std::max(n, 1) * RT_NODE_F4, make_float4(0, 0, 0, 0));
  s.prims.resize((size_t)std::max(n, 1) * RT_PRIM_F4);
  s.ex_prims.resize(std::max(n, 1));
  std::vector<uint32_t> order(n);
  std::iota(order.begin(), order.end(), 0u);
  if (n <= 1) {
    const float inf = std::numeric_limits<float>::infinity();
    for (int a = 0; a < 3; a++) {
      s.nodes[2 * a] = make_float4(n ? boxes[0].lo[a] : inf, inf, inf, inf);
      s.nodes[2 * a + 1] = make_float4(n ? boxes[0].hi[a] : -inf, -inf, -inf, -inf);
    }
    s.nodes[6] = make_float4(ibits(n ? ~0 : RT_EMPTY), ibits(RT_EMPTY), ibits(RT_EMPTY), ibits(RT_EMPTY));
    s.nodes[7] = make_float4(ibits(-1), 0.f, 0.f, 0.f);
    s.n_wide = 1;
  } else {
    float bounds[6];
    for (int a = 0; a < 3; a++) {
      bounds[a] = (float)centroids.lo[a];
      double ext = centroids.hi[a] - centroids.lo[a];
      bounds[3 + a] = ext > 0 ? (float)(1.0 / ext) : 0.f;
    }
    std::vector<BuildBox> sorted_boxes(n);
    std::vector<int> left(n - 1), right(n - 1), parent(2 * n - 1, -2);
    std::vector<BuildBox> box(n - 1);
    std::vector<unsigned int> visits(n - 1, 0);
    // the same choice as rt_scene.cu: host SAH tree for small scenes, a device-style tree (the Karras LBVH, or PLOC
    // on request) otherwise; RT_BVH=best builds the SAH and the PLOC tree and keeps the smaller surface-area sum
    const char *bvh_env = std::getenv("RT_BVH");
    const bool best_of = bvh_env && !std::strcmp(bvh_env, "best") && n >= RT_SAH_MIN_PRIMS && n <= RT_SAH_MAX_PRIMS;
    const bool host_sah = rtsah::use_sah(n);
    const bool device_tree = !host_sah || best_of;
    double host_area = 0.0;
    if (host_sah) {
      rtsah::HostTree ht;
      rtsah::build(boxes.data(), n, ht);
      order = ht.order;
      left = ht.left;
      right = ht.right;
      parent = ht.parent;
      box = ht.box;
      for (const BuildBox &b : box)
        host_area += (double)box_area(b);
    }
    if (device_tree) {
      std::vector<uint32_t> d_order(n);
      std::iota(d_order.begin(), d_order.end(), 0u);
      std::vector<int> d_left(n - 1), d_right(n - 1), d_parent(2 * n - 1, -2);
      std::vector<BuildBox> d_box(n - 1), d_sorted(n);
      std::vector<uint64_t> codes(n), sorted_codes(n);
      for (int i = 0; i < n; i++)
        codes[i] = morton_body(boxes[i], bounds, bounds + 3);
      std::stable_sort(d_order.begin(), d_order.end(), [&](uint32_t a, uint32_t b) { return codes[a] < codes[b]; });
      for (int j = 0; j < n; j++) {
        sorted_codes[j] = codes[d_order[j]];
        d_sorted[j] = boxes[d_order[j]];
      }
      BinTree t{d_left.data(), d_right.data(), d_parent.data(), d_box.data(), visits.data(), n};
      if (rtsah::use_ploc(n) || best_of) { // k_ploc_nearest / scan / k_ploc_merge, one round per iteration
        std::vector<PlocCluster> clusters(n), next(n);
        for (int j = 0; j < n; j++)
          clusters[j] = PlocCluster{d_sorted[j], ~j, 0};
        std::vector<int> nearest(n);
        int count = n, next_node = n - 2;
        while (count > 1) {
          for (int i = 0; i < count; i++)
            nearest[i] = ploc_nearest_body(clusters.data(), count, i);
          int slot = 0, merges = 0;
          for (int i = 0; i < count; i++) {
            int role = ploc_role(nearest.data(), i);
            ploc_merge_body(clusters.data(), nearest.data(), i, role, slot, next_node - merges, next.data(), t);
            slot += role >= 0;
            merges += role > 0;
          }
          next_node -= merges;
          count = slot;
          clusters.swap(next);
        }
      } else {
        for (int i = 0; i < n - 1; i++)
          hierarchy_body(sorted_codes.data(), t, i);
        for (int j = 0; j < n; j++) { // k_refit
          int node = d_parent[(n - 1) + j];
          while (node >= 0) {
            if (visits[node]++ == 0)
              break;
            d_box[node] = box_union(child_box(t, d_sorted.data(), d_left[node]), child_box(t, d_sorted.data(), d_right[node]));
            node = d_parent[node];
          }
        }
      }
      double device_area = 0.0;
      for (const BuildBox &b : d_box)
        device_area += (double)box_area(b);
      if (!host_sah || device_area < host_area) {
        order = d_order;
        left = d_left;
        right = d_right;
        parent = d_parent;
        box = d_box;
      }
    }
    for (int j = 0; j < n; j++)
      sorted_boxes[j] = boxes[order[j]];
    BinTree t{left.data(), right.data(), parent.data(), box.data(), visits.data(), n};
    std::vector<CollapseItem> items{{0, 0, -1}}, next;
    int wide_count = 1;
    while (!items.empty()) { // k_collapse, one level per iteration
      next.clear();
      for (const CollapseItem &it : items) {
        int child[4];
        int n_child = collapse_gather(t, sorted_boxes.data(), it.bin, child);
        int wide_ref[4] = {0, 0, 0, 0};
        for (int k = 0; k < n_child; k++)
          if (child[k] >= 0) {
            wide_ref[k] = wide_count++;
            next.push_back({child[k], wide_ref[k], it.wide * 4 + k});
          }
        collapse_write(t, sorted_boxes.data(), s.nodes.data(), it.wide, child, n_child, wide_ref, it.up);
      }
      items.swap(next);
    }
    s.n_wide = wide_count;
  }
  { // rt_scene.cu: where each surface sphere of the description ended up
    std::vector<int> leaf_of_record(std::max(n, 1), -1);
    for (int j = 0; j < n; j++)
      leaf_of_record[order[j]] = j;
    s.sphere_leaf.assign(s.desc->n_spheres, -1);
    int record = 0;
    for (int i = 0; i < s.desc->n_spheres; i++)
      if (!(s.desc->spheres[i].flags & RT_PRIM_BOUNDARY))
        s.sphere_leaf[i] = leaf_of_record[record++];
    s.quad_leaf.assign(s.desc->n_quads, -1);
    for (int i = 0; i < s.desc->n_quads; i++)
      if (!(s.desc->quads[i].flags & RT_PRIM_BOUNDARY))
        s.quad_leaf[i] = leaf_of_record[record++];
  }
  s.leaf_object.assign(std::max(n, 1), -1);
  s.leaf_id.assign(std::max(n, 1), -1);
  for (int j = 0; j < n; j++) {
    for (int k = 0; k < RT_PRIM_F4; k++)
      s.prims[(size_t)j * RT_PRIM_F4 + k] = f.prims[(size_t)order[j] * RT_PRIM_F4 + k];
    s.ex_prims[j] = f.ex_prims[order[j]];
    s.leaf_object[j] = s.ex_prims[j].object;
    s.leaf_id[j] = s.ex_prims[j].id;
  }
}

extern "C" {

const char *emu_last_error(void) { return g_err.c_str(); }

EmuScene *emu_scene_create(const rt_scene_desc *desc) {
  EmuScene *s = new EmuScene();
  if (rtflat::flatten(desc, s->flat) != RT_OK) {
    delete s;
    return nullptr;
  }
  s->desc = desc;
  build_bvh(*s);
  rtflat::Flat &f = s->flat;
  static const float4 zero4 = {0, 0, 0, 0};
  static const unsigned char zero1 = 0;
  s->d.nodes = s->nodes.data();
  s->d.prims = s->prims.data();
  s->d.bprims = f.bprims.empty() ? &zero4 : f.bprims.data();
  s->d.mats = f.mats.empty() ? &zero4 : f.mats.data();
  s->d.lights = f.lights.empty() ? &zero4 : f.lights.data();
  s->d.perlin_grad = f.perlin_grad.empty() ? &zero4 : f.perlin_grad.data();
  s->d.perlin_perm = f.perlin_perm.empty() ? &zero1 : f.perlin_perm.data();
  static const uint32_t zero_texel = 0;
  s->d.texels = f.texels.empty() ? &zero_texel : f.texels.data();
  s->d.n_prims = (int)f.boxes.size();
  s->d.n_lights = desc->n_lights;
  s->d.n_media = desc->n_media;
  s->ex.nodes = s->nodes.data();
  s->ex.prims = s->ex_prims.data();
  s->ex.bprims = f.ex_bprims.data();
  s->ex.ops = f.ops.data();
  s->ex.chain_first = f.chain_first.data();
  s->ex.chain_count = f.chain_count.data();
  return s;
}

void emu_scene_destroy(EmuScene *s) { delete s; }

// rt_scene_update_spheres_impl (rt_scene.cu) with k_leaf_links / k_update_leaves / k_refit_wide as serial loops.
static int emu_update(EmuScene *s, int first, int count, const rt_sphere *spheres, const rt_quad *quads);
int emu_scene_update_spheres(EmuScene *s, int first, int count, const rt_sphere *spheres) {
  return emu_update(s, first, count, spheres, nullptr);
}
int emu_scene_update_quads(EmuScene *s, int first, int count, const rt_quad *quads) {
  return emu_update(s, first, count, nullptr, quads);
}
static int emu_update(EmuScene *s, int first, int count, const rt_sphere *spheres, const rt_quad *quads) {
  using namespace rtflat;
  Baker bk{s->desc};
  const int n_leaf = s->d.n_prims;
  std::vector<int> leaf_up(std::max(n_leaf, 1), -1);
  for (int node = 0; node < s->n_wide; node++)
    leaf_links_body(s->nodes.data(), node, leaf_up.data());
  for (int k = 0; k < count; k++) {
    int leaf = spheres ? s->sphere_leaf[first + k] : s->quad_leaf[first + k];
    if (leaf < 0)
      return 1;
    std::vector<float4> rec;
    std::vector<PrimExact> ex;
    BoxD box;
    if (spheres) {
      push_sphere(bk, spheres[k], first + k, spheres[k].material, rec, ex, box);
      embed_sphere_material(rec.data(), s->flat.mats);
    } else {
      push_quad(bk, quads[k], s->desc->n_spheres + first + k, quads[k].material, rec, ex, box);
    }
    for (int q = 0; q < RT_PRIM_F4; q++)
      s->prims[(size_t)leaf * RT_PRIM_F4 + q] = rec[q];
    s->ex_prims[leaf] = ex[0];
    node_set_slot_box(s->nodes.data(), leaf_up[leaf] >> 2, leaf_up[leaf] & 3, to_build_box(box));
  }
  std::vector<unsigned int> arrivals(std::max(s->n_wide, 1), 0);
  for (int j = 0; j < n_leaf; j++) {
    int node = leaf_up[j] >> 2;
    for (;;) {
      if (++arrivals[node] < (unsigned int)node_child_count(s->nodes.data(), node))
        break;
      const float4 *n = s->nodes.data() + (size_t)node * RT_NODE_F4;
      int up = f2i(n[7].x);
      if (up < 0)
        break;
      node_set_slot_box(s->nodes.data(), up >> 2, up & 3, node_bounds(n, n[6]));
      node = up >> 2;
    }
  }
  return 0;
}
void emu_stats(uint64_t *nodes, uint64_t *leaves, int reset) {
  *nodes = g_stat_nodes;
  *leaves = g_stat_leaves;
  if (reset)
    g_stat_nodes = g_stat_leaves = 0;
}
int emu_scene_nodes(const EmuScene *s) { return s->n_wide; }
int emu_scene_leaves(const EmuScene *s) { return s->d.n_prims; }

// Surface-area cost of the BVH4: sum over the wide nodes of area(node) / area(root) - the expected number of node
// visits of a random ray that hits the root box (the quantity rt_scene.cu compares when it has two trees).
double emu_scene_cost(const EmuScene *s) {
  double sum = 0.0, root = 0.0;
  for (int i = 0; i < s->n_wide; i++) {
    const float4 *n = s->nodes.data() + (size_t)i * RT_NODE_F4;
    float4 rows[6] = {n[0], n[1], n[2], n[3], n[4], n[5]};
    double a = box_area(node_bounds(rows, n[6]));
    if (i == 0)
      root = a;
    sum += a;
  }
  return root > 0 ? sum / root : 0.0;
}

// Structural check of the BVH4: every leaf referenced exactly once, every child box inside its
// parent's slot box.  Returns 0 when consistent, otherwise a small error code.
int emu_scene_check_bvh(const EmuScene *s) {
  int n = s->d.n_prims;
  std::vector<int> seen(std::max(n, 1), 0);
  std::vector<int> stack{0};
  std::vector<BuildBox> bound_stack;
  BuildBox world;
  for (int a = 0; a < 3; a++) {
    world.lo[a] = -std::numeric_limits<float>::infinity();
    world.hi[a] = std::numeric_limits<float>::infinity();
  }
  bound_stack.push_back(world);
  int visited_nodes = 0;
  while (!stack.empty()) {
    int node = stack.back();
    stack.pop_back();
    BuildBox bound = bound_stack.back();
    bound_stack.pop_back();
    if (node < 0 || node >= s->n_wide)
      return 1;
    visited_nodes++;
    const float4 *nd = s->nodes.data() + (size_t)node * RT_NODE_F4;
    const float *lo[3] = {&nd[0].x, &nd[2].x, &nd[4].x}, *hi[3] = {&nd[1].x, &nd[3].x, &nd[5].x};
    const float *refs = &nd[6].x;
    for (int c = 0; c < 4; c++) {
      int ref = f2i(refs[c]);
      if (ref == RT_EMPTY)
        continue;
      BuildBox b;
      for (int a = 0; a < 3; a++) {
        b.lo[a] = lo[a][c];
        b.hi[a] = hi[a][c];
        if (b.lo[a] < bound.lo[a] || b.hi[a] > bound.hi[a])
          return 2;
      }
      if (ref < 0) {
        int leaf = ~ref;
        if (leaf >= n)
          return 3;
        seen[leaf]++;
      } else {
        stack.push_back(ref);
        bound_stack.push_back(b);
      }
    }
  }
  for (int j = 0; j < n; j++)
    if (seen[j] != 1)
      return 4;
  if (visited_nodes != s->n_wide)
    return 5;
  return 0;
}

void emu_trace(EmuScene *s, const rt_ray *rays, int64_t n, int mode, uint64_t seed, rt_hit *hits) {
  for (int64_t q = 0; q < n; q++) {
    const rt_ray &in = rays[q];
    RayKey key{seed, in.rng_pixel, in.rng_sample, in.rng_bounce};
    rt_hit out{};
    if (mode == RT_TRACE_EXACT_F64) {
      RayD r;
      for (int k = 0; k < 3; k++) {
        r.o[k] = in.origin[k];
        r.d[k] = in.direction[k];
      }
      r.time = in.time;
      HitD h;
      traverse_exact(s->ex, r, in.t_min, in.t_max, h, key);
      out.t = h.prim >= 0 ? h.t : (double)RT_INF_F;
      out.prim = h.id;
      out.object = h.object;
      out.front_face = h.prim >= 0 ? h.front : 0;
    } else {
      Ray r;
      r.o = F3((float)in.origin[0], (float)in.origin[1], (float)in.origin[2]);
      r.d = F3((float)in.direction[0], (float)in.direction[1], (float)in.direction[2]);
      r.time = (float)in.time;
      Hit best{(float)in.t_max, -1};
      LocalStack stack;
      traverse(s->d, r, (float)in.t_min, best, -1, key, stack);
      out.t = best.prim >= 0 ? (double)best.t : (double)RT_INF_F;
      out.prim = best.prim >= 0 ? s->leaf_id[best.prim] : -1;
      out.object = best.prim >= 0 ? s->leaf_object[best.prim] : -1;
      out.front_face = 0;
    }
    hits[q] = out;
  }
}

// The wavefront pass of rt_api.cu / rt_kernels.cu as serial loops: generate -> (extend, shade)* ->
// accumulate, for the full image (rank 0 of 1).  out_rgb = sum over the strata * scale.
void emu_render(EmuScene *s, const rt_camera *camera, int first_sample, int n_samples, int sqrt_spp, int max_depth,
                uint64_t seed, double scale, float *out_rgb, uint64_t *segments) {
  DCamera cam{};
  for (int a = 0; a < 3; a++) {
    cam.center[a] = (float)camera->center[a];
    cam.p00c[a] = (float)(camera->pixel00_loc[a] - camera->center[a]);
    cam.du[a] = (float)camera->pixel_delta_u[a];
    cam.dv[a] = (float)camera->pixel_delta_v[a];
    cam.disk_u[a] = (float)camera->defocus_disk_u[a];
    cam.disk_v[a] = (float)camera->defocus_disk_v[a];
    s->d.bg[a] = (float)camera->background[a];
  }
  cam.defocus = camera->defocus_angle > 0;
  cam.width = camera->image_width;
  cam.height = camera->image_height;
  const int W = cam.width, H = cam.height, npix = W * H;
  uint64_t segs = 0;
  std::vector<float> film((size_t)npix * 3, 0.f);
  struct Q {
    Ray r;
    int path, skip;
  };
  for (int smp = 0; smp < n_samples; smp++) {
    int sidx = first_sample + smp;
    std::vector<Q> queue(npix), next;
    std::vector<f3> throughput(npix, F3(1.f, 1.f, 1.f)), radiance(npix, F3(0.f, 0.f, 0.f));
    for (int p = 0; p < npix; p++) { // k_generate
      int row = p / W, col = p % W;
      Uniform4 u0 = philox_uniform4(seed, (uint32_t)p, (uint32_t)sidx, 0, RT_STREAM_CAMERA, 0);
      Uniform4 u1 = philox_uniform4(seed, (uint32_t)p, (uint32_t)sidx, 0, RT_STREAM_CAMERA, 1);
      queue[p].r = camera_ray(cam, col, row, sidx % sqrt_spp, sidx / sqrt_spp, (float)(1.0 / sqrt_spp), u0, u1);
      queue[p].path = p;
      queue[p].skip = -1;
    }
    for (int bounce = 0; bounce < max_depth && !queue.empty(); bounce++) {
      next.clear();
      segs += queue.size();
      for (const Q &q : queue) {
        RayKey key{seed, (uint32_t)q.path, (uint32_t)sidx, (uint32_t)bounce};
        Hit best{RT_INF_F, -1};
        LocalStack stack;
        traverse(s->d, q.r, RT_T_MIN, best, q.skip, key, stack); // k_extend
        ShadeResult res;
        bool cont = shade_segment(s->d, q.r, best, throughput[q.path], key, bounce + 1 >= max_depth, res); // k_shade
        if (cont) {
          throughput[q.path] = res.throughput;
          next.push_back({res.next, q.path, res.next_skip_prim});
        } else {
          radiance[q.path] = res.radiance;
        }
      }
      queue.swap(next);
    }
    for (int p = 0; p < npix; p++) { // k_accumulate
      film[(size_t)p * 3 + 0] += radiance[p].x;
      film[(size_t)p * 3 + 1] += radiance[p].y;
      film[(size_t)p * 3 + 2] += radiance[p].z;
    }
  }
  for (size_t k = 0; k < film.size(); k++)
    out_rgb[k] = (float)(scale * (double)film[k]);
  if (segments)
    *segments = segs;
}

// PathMap (rt_device.h): fills out[4 * i ..] = (sample in pass, film index, global scanline, column) for paths
// first .. first + count - 1 of a pass over the tiles of `rank`; returns whether the 8 x 4 block order is in use
int emu_path_map(int width, int height, int rank, int n_ranks, int tile_rows, int allow_blocks, uint32_t first,
                 uint32_t count, uint32_t *out) {
  DFilmMap map{width, height, rank, n_ranks, tile_rows};
  long long n_owned = (long long)owned_rows(height, rank, n_ranks, tile_rows) * width;
  PathMap m = pathmap_make(map, n_owned, allow_blocks != 0);
  for (uint32_t i = 0; i < count; i++) {
    uint32_t s, k;
    int row, col;
    path_to_pixel(m, first + i, s, k, row, col);
    out[4 * i + 0] = s;
    out[4 * i + 1] = k;
    out[4 * i + 2] = (uint32_t)row;
    out[4 * i + 3] = (uint32_t)col;
  }
  return m.tiled;
}

// The FP32 slab test of node_visit (rt_device.h) on one box: returns 1 when the box would be entered.
// bias[a] = -1 / 0 / +1 moves 1 / d of axis a by that many ulps off the host's correctly rounded quotient: the device
// takes the reciprocal with one instruction (error <= 2^-23 relative), and the guarantee has to hold for that too.
int emu_box_test_biased(const float *lo, const float *hi, const float *o, const float *d, float tmin, float tmax,
                        const int *bias) {
  float4 node[RT_NODE_F4];
  const float inf = RT_INF_F;
  for (int a = 0; a < 3; a++) {
    node[2 * a] = make_float4(lo[a], inf, inf, inf);
    node[2 * a + 1] = make_float4(hi[a], -inf, -inf, -inf);
  }
  node[6] = make_float4(i2f(~0), i2f(RT_EMPTY), i2f(RT_EMPTY), i2f(RT_EMPTY));
  node[7] = make_float4(i2f(-1), 0.f, 0.f, 0.f);
  DScene sc{};
  sc.nodes = node;
  RayTrav rt = make_trav(F3(o[0], o[1], o[2]), F3(d[0], d[1], d[2]));
  if (bias) {
    float inv[3] = {rt.inv.x, rt.inv.y, rt.inv.z};
    for (int a = 0; a < 3; a++)
      for (int k = 0; k < (bias[a] < 0 ? -bias[a] : bias[a]); k++)
        inv[a] = std::nextafterf(inv[a], bias[a] < 0 ? -inf : inf);
    rt = trav_from(F3(inv[0], inv[1], inv[2]), F3(o[0] * inv[0], o[1] * inv[1], o[2] * inv[2]));
  }
  LocalStack stack;
  int sp = 0, next = 0;
  return node_visit(sc, 0, rt, tmin, tmax, stack, sp, next) ? 1 : 0;
}
int emu_box_test(const float *lo, const float *hi, const float *o, const float *d, float tmin, float tmax) {
  return emu_box_test_biased(lo, hi, o, d, tmin, tmax, nullptr);
}

// sphere_hit (rt_device.h) on one static sphere: returns 1 and *t on a hit
int emu_sphere_test(const float *center, float radius, const float *o, const float *d, float tmin, float tmax, float *t) {
  Ray r;
  r.o = F3(o[0], o[1], o[2]);
  r.d = F3(d[0], d[1], d[2]);
  r.time = 0.f;
  return sphere_hit(make_float4(center[0], center[1], center[2], radius), make_float4(0.f, 0.f, 0.f, 0.f), r, tmin, tmax, *t) ? 1 : 0;
}

// quad_hit (rt_device.h) on one quad given as the reference's Plane(Q, u, v): record made by rt_flatten.h
int emu_quad_test(const double *corner, const double *u, const double *v, const float *o, const float *d, float tmin,
                  float tmax, float *t) {
  rt_quad q{};
  for (int k = 0; k < 3; k++)
    q.corner[k] = corner[k], q.u[k] = u[k], q.v[k] = v[k];
  q.xform = -1;
  rtflat::Baker bk{nullptr};
  std::vector<float4> rec;
  std::vector<PrimExact> ex;
  rtflat::BoxD box;
  rtflat::push_quad(bk, q, 0, 0, rec, ex, box);
  Ray r;
  r.o = F3(o[0], o[1], o[2]);
  r.d = F3(d[0], d[1], d[2]);
  r.time = 0.f;
  return quad_hit(rec[0], rec[1], rec[2], rec[3].x, r, tmin, tmax, *t) ? 1 : 0;
}

// x / d through the device code's FastDiv
uint32_t emu_fastdiv(uint32_t d, uint32_t x) { return fastdiv(fastdiv_make(d), x); }

} // extern "C"
