// rt_bvh.h — per-thread bodies of the GPU BVH build (Morton codes -> radix sort -> binary tree over the sorted
// primitives: Karras hierarchy + bottom-up refit [LBVH], or locally-ordered clustering [PLOC] -> collapse to
// 4-wide SoA nodes).  Replaces the reference's host-side recursive
// SAH build and its pointer / 64-byte FlatNode tree (optimization/BVHNode.cpp:21-123,322-383).
//
// Stages (one kernel each in rt_kernels.cu, thread i runs body(i)):
//   morton_body      63-bit Morton code of the primitive box centre inside the scene bounds
//   (sort)           cub::DeviceRadixSort on (code, primitive index)
//   hierarchy_body   Karras 2012: internal node i of the binary radix tree over the sorted codes
//   refit_body       leaf i climbs to the root; the second thread to reach a node merges the boxes
//   collapse_body    binary node -> 4-wide node: repeatedly open the child with the largest surface
//                    area until four children; work items for the next level go to a queue
#pragma once

#include "rt_device.h"

struct BuildBox {
  float lo[3], hi[3];
};

// Binary radix tree.  Node ids: internal i in [0, n-1), leaf j is encoded as ~j.
struct BinTree {
  int *left, *right, *parent; // per internal node; parent also per leaf at offset n_internal
  BuildBox *box;              // per internal node
  unsigned int *visits;       // refit arrival counters
  int n;                      // number of leaves
};

RT_HD uint64_t expand_bits21(uint32_t v) {
  uint64_t x = v & 0x1fffffu;
  x = (x | x << 32) & 0x1f00000000ffffull;
  x = (x | x << 16) & 0x1f0000ff0000ffull;
  x = (x | x << 8) & 0x100f00f00f00f00full;
  x = (x | x << 4) & 0x10c30c30c30c30c3ull;
  x = (x | x << 2) & 0x1249249249249249ull;
  return x;
}

RT_HD uint64_t morton_body(const BuildBox &b, const float *scene_lo, const float *scene_inv_extent) {
  uint32_t q[3];
  for (int a = 0; a < 3; a++) {
    float c = 0.5f * (b.lo[a] + b.hi[a]);
    float u = (c - scene_lo[a]) * scene_inv_extent[a];
    u = fminf(fmaxf(u, 0.f), 1.f);
    uint32_t v = (uint32_t)(u * 2097151.0f);
    q[a] = v > 2097151u ? 2097151u : v;
  }
  return (expand_bits21(q[0]) << 2) | (expand_bits21(q[1]) << 1) | expand_bits21(q[2]);
}

RT_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
RT_HD int clz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

// Common-prefix length of sorted keys i and j, index as tie-break for equal codes; -1 out of range.
RT_HD int prefix_delta(const uint64_t *codes, int n, int i, int j) {
  if (j < 0 || j >= n)
    return -1;
  uint64_t a = codes[i], b = codes[j];
  if (a == b)
    return 64 + clz32((uint32_t)i ^ (uint32_t)j);
  return clz64(a ^ b);
}

RT_HD void hierarchy_body(const uint64_t *codes, BinTree t, int i) {
  int n = t.n;
  int d = (prefix_delta(codes, n, i, i + 1) - prefix_delta(codes, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = prefix_delta(codes, n, i, i - d);
  int lmax = 2;
  while (prefix_delta(codes, n, i, i + lmax * d) > dmin)
    lmax *= 2;
  int l = 0;
  for (int s = lmax / 2; s >= 1; s /= 2)
    if (prefix_delta(codes, n, i, i + (l + s) * d) > dmin)
      l += s;
  int j = i + l * d;
  int dnode = prefix_delta(codes, n, i, j);
  int s = 0;
  int div = 2;
  for (;;) {
    int step = (l + div - 1) / div;
    if (prefix_delta(codes, n, i, i + (s + step) * d) > dnode)
      s += step;
    if (step <= 1)
      break;
    div *= 2;
  }
  int gamma = i + s * d + (d < 0 ? -1 : 0);
  int lo = i < j ? i : j, hi = i < j ? j : i;
  int left = (lo == gamma) ? ~gamma : gamma;
  int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
  t.left[i] = left;
  t.right[i] = right;
  int n_internal = n - 1;
  if (left >= 0)
    t.parent[left] = i;
  else
    t.parent[n_internal + ~left] = i;
  if (right >= 0)
    t.parent[right] = i;
  else
    t.parent[n_internal + ~right] = i;
  if (i == 0)
    t.parent[0] = -1;
}

RT_HD BuildBox box_union(const BuildBox &a, const BuildBox &b) {
  BuildBox r;
  for (int k = 0; k < 3; k++) {
    r.lo[k] = fminf(a.lo[k], b.lo[k]);
    r.hi[k] = fmaxf(a.hi[k], b.hi[k]);
  }
  return r;
}
RT_HD float box_area(const BuildBox &b) {
  float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  return 2.f * (dx * dy + dy * dz + dz * dx);
}

RT_HD BuildBox child_box(const BinTree &t, const BuildBox *leaf_boxes, int ref) {
  return ref >= 0 ? t.box[ref] : leaf_boxes[~ref];
}

// ---------------------------------------------------------------------------------------------------
// PLOC (parallel locally-ordered clustering; Meister & Bittner 2018): an agglomerative build over the
// Morton-sorted primitives.  The clusters of the current level keep the Morton order; every cluster looks at
// its `radius` neighbours on either side and picks the one whose merged box has the smallest surface area;
// clusters that picked each other merge into a new binary node, the others survive to the next round; the
// array is compacted and the rounds repeat until one cluster is left (~2 log2 n rounds).  Same input and
// output as the Karras hierarchy + refit pair (sorted leaf boxes in, BinTree out, root = node 0) at a
// surface-area cost close to a top-down SAH build, which is what traversal cost follows.
// ---------------------------------------------------------------------------------------------------
#define RT_PLOC_RADIUS 16

struct PlocCluster {
  BuildBox box;
  int ref; // binary-tree reference: internal node index, or ~(sorted leaf index)
  int pad_;
};

// Nearest neighbour of cluster i among clusters [i - radius, i + radius] by merged surface area (ties: the
// lower index, so that the choice does not depend on the evaluation order).
RT_HD int ploc_nearest_body(const PlocCluster *clusters, int count, int i) {
  const BuildBox me = clusters[i].box;
  int lo = i - RT_PLOC_RADIUS < 0 ? 0 : i - RT_PLOC_RADIUS;
  int hi = i + RT_PLOC_RADIUS > count - 1 ? count - 1 : i + RT_PLOC_RADIUS;
  int best = -1;
  float best_area = RT_INF_F;
  for (int j = lo; j <= hi; j++) {
    if (j == i)
      continue;
    float a = box_area(box_union(me, clusters[j].box));
    if (a < best_area) {
      best_area = a;
      best = j;
    }
  }
  return best;
}

// What cluster i does this round: 1 = merges with nearest[i] and creates the node (the lower index of a mutual
// pair), -1 = is absorbed by its partner, 0 = survives unchanged.
RT_HD int ploc_role(const int *nearest, int i) {
  int j = nearest[i];
  if (j < 0 || nearest[j] != i)
    return 0;
  return i < j ? 1 : -1;
}

// Writes cluster i's successor into next[slot] (slot = number of surviving / merging clusters before i) and, for
// a merging cluster, the new binary node `node` (= first_node - number of merging clusters before i: nodes are
// handed out downwards from n - 2, so the last merge of the build creates node 0, the root).
RT_HD void ploc_merge_body(const PlocCluster *clusters, const int *nearest, int i, int role, int slot, int node,
                           PlocCluster *next, BinTree t) {
  if (role < 0)
    return;
  PlocCluster c = clusters[i];
  if (role > 0) {
    const PlocCluster &o = clusters[nearest[i]];
    const int n_internal = t.n - 1;
    t.left[node] = c.ref;
    t.right[node] = o.ref;
    c.box = box_union(c.box, o.box);
    t.box[node] = c.box;
    t.parent[c.ref >= 0 ? c.ref : n_internal + ~c.ref] = node;
    t.parent[o.ref >= 0 ? o.ref : n_internal + ~o.ref] = node;
    if (node == 0)
      t.parent[0] = -1;
    c.ref = node;
  }
  next[slot] = c;
}

// Collapse work item: binary node `bin` becomes wide node `wide`.
struct CollapseItem {
  int bin, wide;
  int up; // parent wide node * 4 + slot in it, -1 for the root: stored in the node's spare row for refits
};

// Returns the number of internal children; their work items are written to next[*] by the caller
// through `alloc` (index of the first of `count` consecutive new wide nodes).
RT_HD int collapse_gather(const BinTree &t, const BuildBox *leaf_boxes, int bin, int child[4]) {
  int n = 2;
  child[0] = t.left[bin];
  child[1] = t.right[bin];
  while (n < 4) {
    int best = -1;
    float best_area = -1.f;
    for (int k = 0; k < n; k++)
      if (child[k] >= 0) {
        float a = box_area(t.box[child[k]]);
        if (a > best_area) {
          best_area = a;
          best = k;
        }
      }
    if (best < 0)
      break;
    int open = child[best];
    child[best] = t.left[open];
    child[n++] = t.right[open];
  }
  (void)leaf_boxes;
  return n;
}

RT_HD void collapse_write(const BinTree &t, const BuildBox *leaf_boxes, float4 *nodes, int wide, const int child[4],
                          int n_child, const int wide_ref[4], int up) {
  float lo[3][4], hi[3][4];
  int ref[4];
  for (int k = 0; k < 4; k++) {
    if (k < n_child) {
      BuildBox b = child_box(t, leaf_boxes, child[k]);
      for (int a = 0; a < 3; a++) {
        lo[a][k] = b.lo[a];
        hi[a][k] = b.hi[a];
      }
      ref[k] = child[k] >= 0 ? wide_ref[k] : child[k]; // leaf refs keep their ~(sorted index) encoding
    } else {
      for (int a = 0; a < 3; a++) {
        lo[a][k] = RT_INF_F;
        hi[a][k] = -RT_INF_F;
      }
      ref[k] = RT_EMPTY;
    }
  }
  float4 *n = nodes + (size_t)wide * RT_NODE_F4;
  n[0] = make_float4(lo[0][0], lo[0][1], lo[0][2], lo[0][3]);
  n[1] = make_float4(hi[0][0], hi[0][1], hi[0][2], hi[0][3]);
  n[2] = make_float4(lo[1][0], lo[1][1], lo[1][2], lo[1][3]);
  n[3] = make_float4(hi[1][0], hi[1][1], hi[1][2], hi[1][3]);
  n[4] = make_float4(lo[2][0], lo[2][1], lo[2][2], lo[2][3]);
  n[5] = make_float4(hi[2][0], hi[2][1], hi[2][2], hi[2][3]);
  n[6] = make_float4(i2f(ref[0]), i2f(ref[1]), i2f(ref[2]), i2f(ref[3]));
  n[7] = make_float4(i2f(up), 0.f, 0.f, 0.f);
}

// ---------------------------------------------------------------------------------------------------
// Refit of the finished BVH4 (animated scenes, rt_scene_update_spheres): every node stores its children's
// boxes, so a node's own box is the union of its used slots and lives in its parent's slot; row 7 of a node
// holds `parent * 4 + slot` (-1 at the root).  Leaves climb, the last child to arrive at a node moves on.
// ---------------------------------------------------------------------------------------------------
RT_HD void node_set_slot_box(float4 *nodes, int node, int slot, const BuildBox &b) {
  float *n = reinterpret_cast<float *>(nodes + (size_t)node * RT_NODE_F4);
  for (int a = 0; a < 3; a++) {
    n[(2 * a) * 4 + slot] = b.lo[a];
    n[(2 * a + 1) * 4 + slot] = b.hi[a];
  }
}
RT_HD int node_child_count(const float4 *nodes, int node) {
  const float4 cr = nodes[(size_t)node * RT_NODE_F4 + 6];
  return (f2i(cr.x) != RT_EMPTY) + (f2i(cr.y) != RT_EMPTY) + (f2i(cr.z) != RT_EMPTY) + (f2i(cr.w) != RT_EMPTY);
}
// rows[0..5] = the node's six box rows as the caller read them (the device reads them past L1)
RT_HD BuildBox node_bounds(const float4 rows[6], const float4 cr) {
  const int ref[4] = {f2i(cr.x), f2i(cr.y), f2i(cr.z), f2i(cr.w)};
  BuildBox b;
  for (int a = 0; a < 3; a++) {
    const float lo[4] = {rows[2 * a].x, rows[2 * a].y, rows[2 * a].z, rows[2 * a].w};
    const float hi[4] = {rows[2 * a + 1].x, rows[2 * a + 1].y, rows[2 * a + 1].z, rows[2 * a + 1].w};
    b.lo[a] = RT_INF_F;
    b.hi[a] = -RT_INF_F;
    for (int k = 0; k < 4; k++)
      if (ref[k] != RT_EMPTY) {
        b.lo[a] = fminf(b.lo[a], lo[k]);
        b.hi[a] = fmaxf(b.hi[a], hi[k]);
      }
  }
  return b;
}
// leaf -> (node * 4 + slot) of the slot that references it; thread `node` fills the entries of its leaf children
RT_HD void leaf_links_body(const float4 *nodes, int node, int *leaf_up) {
  const float4 cr = nodes[(size_t)node * RT_NODE_F4 + 6];
  const int ref[4] = {f2i(cr.x), f2i(cr.y), f2i(cr.z), f2i(cr.w)};
  for (int k = 0; k < 4; k++)
    if (ref[k] < 0)
      leaf_up[~ref[k]] = node * 4 + k;
}
