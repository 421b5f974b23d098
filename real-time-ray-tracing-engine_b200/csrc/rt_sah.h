// rt_sah.h — host-side binary SAH build for small scenes (the same cost model as the reference's builder,
// optimization/BVHNode.cpp:168-320: surface-area heuristic over candidate planes on the centroid bounds, median
// fallback), written as an iterative build: 32 bins per axis on large nodes, an exact sweep on nodes of <= 16.  Produces the BinTree arrays the device collapse stage
// (rt_bvh.h, k_collapse) turns into 4-wide nodes, so everything downstream of the binary tree is shared with the
// GPU LBVH path.  Used from RT_SAH_MIN_PRIMS to RT_SAH_MAX_PRIMS primitives (0.7 ms at 485, 120 ms at 65,536 on one
// host thread), where the better tree saves 11-39 % of the node visits of every ray; larger scenes keep the
// device LBVH (2 ms for 10^6 primitives).
#pragma once

#include "rt_bvh.h"

#include <algorithm>
#include <cstdint>
#include <vector>

#include <cstdlib>
#include <cstring>

#define RT_SAH_MIN_PRIMS 64 // below this the tree is two levels either way (measured: no gain, Cornell + smoke loses)
#define RT_SAH_MAX_PRIMS (1 << 16)

namespace rtsah {

// Set (per thread) by rt_scene_build while it rebuilds a scene whose device tree came out deeper than the traversal
// stack allows: the host SAH builder splits at the spatial / object median when its sweep finds nothing better, so its
// trees stay shallow where a radix tree over clustered Morton codes does not.
inline thread_local bool g_force_sah = false;

// RT_BVH=lbvh / ploc / sah force one builder (A/B runs and the tests of every path).
inline bool use_sah(int n) {
  if (g_force_sah)
    return n >= 2;
  const char *e = std::getenv("RT_BVH");
  if (e && (!std::strcmp(e, "lbvh") || !std::strcmp(e, "ploc")))
    return false;
  if (e && !std::strcmp(e, "sah"))
    return n >= 2;
  return n >= RT_SAH_MIN_PRIMS && n <= RT_SAH_MAX_PRIMS;
}
// Device builds: the Karras radix tree by default, PLOC (rt_bvh.h) on request (RT_BVH=ploc).  Measured on B200
// (profiles/r02_experiments.md): PLOC cuts the node visits of the final scene from 6.8 (host SAH) to 5.8 per
// segment and still renders 3 % slower (more of the expensive medium leaves are reached), and on the uniform
// sphere fields it visits 7 % MORE nodes than the radix tree, whose Morton splits are spatial medians of the grid.
inline bool use_ploc(int n) {
  if (g_force_sah)
    return false;
  const char *e = std::getenv("RT_BVH");
  return e && !std::strcmp(e, "ploc") && n >= 2;
}

struct HostTree {
  std::vector<int> left, right, parent; // BinTree layout: parent has 2n-1 entries (internal, then leaves)
  std::vector<BuildBox> box;            // per internal node
  std::vector<uint32_t> order;          // leaf j holds primitive order[j]
};

inline BuildBox empty_box() {
  BuildBox b;
  for (int a = 0; a < 3; a++) {
    b.lo[a] = RT_INF_F;
    b.hi[a] = -RT_INF_F;
  }
  return b;
}
inline float safe_area(const BuildBox &b) { return b.hi[0] < b.lo[0] ? 0.f : box_area(b); }
// box_union without fminf / fmaxf calls (no NaNs in build boxes): the inner loop of the build
inline BuildBox merge(const BuildBox &a, const BuildBox &b) {
  BuildBox r;
  for (int k = 0; k < 3; k++) {
    r.lo[k] = a.lo[k] < b.lo[k] ? a.lo[k] : b.lo[k];
    r.hi[k] = a.hi[k] > b.hi[k] ? a.hi[k] : b.hi[k];
  }
  return r;
}

// boxes: one per primitive (n >= 2).  Leaves hold exactly one primitive.
inline void build(const BuildBox *boxes, int n, HostTree &t) {
  const int n_internal = n - 1;
  t.left.assign(n_internal, 0);
  t.right.assign(n_internal, 0);
  t.parent.assign(2 * (size_t)n - 1, -1);
  t.box.assign(n_internal, empty_box());
  t.order.resize(n);
  for (int i = 0; i < n; i++)
    t.order[i] = (uint32_t)i;
  std::vector<float> cen((size_t)n * 3);
  for (int i = 0; i < n; i++)
    for (int a = 0; a < 3; a++)
      cen[(size_t)i * 3 + a] = 0.5f * (boxes[i].lo[a] + boxes[i].hi[a]);

  struct Task {
    int first, count, node; // range of `order`, internal node index to fill
  };
  std::vector<Task> stack;
  int next_node = 1;
  stack.push_back({0, n, 0});
  constexpr int kBins = 32;
  constexpr int kSweepBelow = 16;
  while (!stack.empty()) {
    Task task = stack.back();
    stack.pop_back();
    uint32_t *idx = t.order.data() + task.first;
    BuildBox bounds = empty_box(), cbounds = empty_box();
    for (int k = 0; k < task.count; k++) {
      bounds = merge(bounds, boxes[idx[k]]);
      for (int a = 0; a < 3; a++) {
        float c = cen[(size_t)idx[k] * 3 + a];
        cbounds.lo[a] = c < cbounds.lo[a] ? c : cbounds.lo[a];
        cbounds.hi[a] = c > cbounds.hi[a] ? c : cbounds.hi[a];
      }
    }
    t.box[task.node] = bounds;

    int mid = -1;
    if (task.count == 2) {
      mid = 1;
    } else if (task.count <= kSweepBelow) {
      // small node: exact sweep over the primitives sorted by centroid, per axis (cheaper than filling bins)
      float best_cost = RT_INF_F;
      int best_axis = -1, best_k = -1;
      uint32_t sorted[3][kSweepBelow];
      for (int a = 0; a < 3; a++) {
        uint32_t *o = sorted[a];
        for (int k = 0; k < task.count; k++) { // insertion sort by (centroid, index)
          uint32_t p = idx[k];
          float cp = cen[(size_t)p * 3 + a];
          int j = k;
          while (j > 0 && (cen[(size_t)o[j - 1] * 3 + a] > cp || (cen[(size_t)o[j - 1] * 3 + a] == cp && o[j - 1] > p))) {
            o[j] = o[j - 1];
            j--;
          }
          o[j] = p;
        }
        float right_area[kSweepBelow];
        BuildBox acc = empty_box();
        for (int k = task.count - 1; k >= 1; k--) {
          acc = merge(acc, boxes[o[k]]);
          right_area[k] = box_area(acc);
        }
        acc = empty_box();
        for (int k = 1; k < task.count; k++) {
          acc = merge(acc, boxes[o[k - 1]]);
          float cost = box_area(acc) * (float)k + right_area[k] * (float)(task.count - k);
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = a;
            best_k = k;
          }
        }
      }
      if (best_axis >= 0) {
        for (int k = 0; k < task.count; k++)
          idx[k] = sorted[best_axis][k];
        mid = best_k;
      } else { // no finite cost (flatten() refuses such geometry; kept as a guard): split in the middle as they lie
        mid = task.count / 2;
      }
    } else {
      // binned SAH over the three axes of the centroid bounds
      float best_cost = RT_INF_F;
      int best_axis = -1, best_bin = -1;
      for (int a = 0; a < 3; a++) {
        float ext = cbounds.hi[a] - cbounds.lo[a];
        if (!(ext > 0.f))
          continue;
        float scale = (float)kBins / ext;
        BuildBox bin_box[kBins];
        int bin_count[kBins];
        for (int b = 0; b < kBins; b++) {
          bin_box[b] = empty_box();
          bin_count[b] = 0;
        }
        for (int k = 0; k < task.count; k++) {
          int b = (int)((cen[(size_t)idx[k] * 3 + a] - cbounds.lo[a]) * scale);
          b = b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
          bin_box[b] = merge(bin_box[b], boxes[idx[k]]);
          bin_count[b]++;
        }
        float right_area[kBins];
        int right_count[kBins];
        BuildBox acc = empty_box();
        int cnt = 0;
        for (int b = kBins - 1; b >= 1; b--) {
          acc = merge(acc, bin_box[b]);
          cnt += bin_count[b];
          right_area[b] = safe_area(acc);
          right_count[b] = cnt;
        }
        acc = empty_box();
        cnt = 0;
        for (int b = 0; b < kBins - 1; b++) { // split between bin b and b + 1
          acc = merge(acc, bin_box[b]);
          cnt += bin_count[b];
          if (cnt == 0 || right_count[b + 1] == 0)
            continue;
          float cost = safe_area(acc) * (float)cnt + right_area[b + 1] * (float)right_count[b + 1];
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = a;
            best_bin = b;
          }
        }
      }
      if (best_axis >= 0) {
        float ext = cbounds.hi[best_axis] - cbounds.lo[best_axis];
        float scale = (float)kBins / ext;
        float lo = cbounds.lo[best_axis];
        uint32_t *m = std::partition(idx, idx + task.count, [&](uint32_t p) {
          int b = (int)((cen[(size_t)p * 3 + best_axis] - lo) * scale);
          b = b < 0 ? 0 : (b > kBins - 1 ? kBins - 1 : b);
          return b <= best_bin;
        });
        mid = (int)(m - idx);
      }
      if (mid <= 0 || mid >= task.count) { // all centroids coincide (or numerical trouble): split in the middle
        mid = task.count / 2;
        int axis = 0;
        for (int a = 1; a < 3; a++)
          if (cbounds.hi[a] - cbounds.lo[a] > cbounds.hi[axis] - cbounds.lo[axis])
            axis = a;
        std::nth_element(idx, idx + mid, idx + task.count, [&](uint32_t p, uint32_t q) {
          return cen[(size_t)p * 3 + axis] < cen[(size_t)q * 3 + axis];
        });
      }
    }
    // children: a single primitive becomes leaf ~position, otherwise a new internal node
    int child[2];
    int first[2] = {task.first, task.first + mid}, count[2] = {mid, task.count - mid};
    for (int c = 0; c < 2; c++) {
      if (count[c] == 1) {
        child[c] = ~first[c];
        t.parent[(size_t)n_internal + first[c]] = task.node;
      } else {
        child[c] = next_node++;
        t.parent[child[c]] = task.node;
        stack.push_back({first[c], count[c], child[c]});
      }
    }
    t.left[task.node] = child[0];
    t.right[task.node] = child[1];
  }
}

} // namespace rtsah
