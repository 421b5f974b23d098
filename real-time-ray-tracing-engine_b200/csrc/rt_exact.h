// rt_exact.h — the FP64 parity path behind rt_trace_rays(RT_TRACE_EXACT_F64).
//
// Same BVH4 as the render path, but the primitive tests are the reference's FP64 formulas in the
// reference's operation order (Sphere.cpp:101-143, Plane.cpp:78-112, Translate.cpp:17-28,
// RotateY.cpp:41-79, ConstantMedium.cpp:25-94), applied in object space through the instance chain
// exactly as the reference's wrapper objects do, so closest-hit ids and t are bit-identical to the
// reference's Hittable::hit.  The translation unit that includes this header is compiled with
// --fmad=false (nvcc) / -ffp-contract=off (host test build): no fused multiply-adds.
// Node boxes are FP32 and padded outward; they are tested in FP64, so the traversal can only visit
// more leaves than the reference, never fewer.
#pragma once

#include "rt_device.h"

struct PrimExact {
  double a[3]; // sphere: center0        quad: corner
  double b[3]; // sphere: center_dir     quad: u
  double c[3]; //                        quad: v
  double n[3]; //                        quad: unit normal
  double w[3]; //                        quad: n / (n.n)
  double s;    // sphere: radius         quad: D           medium: density
  int type;    // RT_PT_*
  int xform;   // instance chain, -1 = none
  int id;      // unified primitive id
  int object;  // top-level world object
  int closed;  // 1: members test t with Interval::contains (quads), 0: surrounds (spheres, media)
  int medium;  // medium index (RT_PT_MEDIUM)
  int first;   // medium: first boundary record
  int count;   // medium: number of boundary records
};

struct XformOpExact {
  int type; // RT_XF_TRANSLATE = 0, RT_XF_ROTATE_Y = 1
  int pad_;
  double offset[3];
  double sin_theta, cos_theta;
};

struct ExactScene {
  const float4 *nodes;
  const PrimExact *prims;  // leaf order
  const PrimExact *bprims; // media boundaries, object space
  const XformOpExact *ops;
  const int *chain_first; // per instance chain
  const int *chain_count;
};

struct RayD {
  double o[3], d[3], time;
};

struct HitD {
  double t;
  int prim; // leaf-order index, -1 = miss
  int front;
  int object;
  int id;
};

RT_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// Ray into the object space of instance chain `xf` (outermost wrapper first).
RT_HD RayD to_object_space(const ExactScene &sc, int xf, const RayD &ray) {
  RayD r = ray;
  if (xf < 0)
    return r;
  int first = sc.chain_first[xf], n = sc.chain_count[xf];
  for (int k = 0; k < n; k++) {
    const XformOpExact &op = sc.ops[first + k];
    if (op.type == 0) {
      r.o[0] = r.o[0] - op.offset[0];
      r.o[1] = r.o[1] - op.offset[1];
      r.o[2] = r.o[2] - op.offset[2];
    } else {
      double c = op.cos_theta, s = op.sin_theta;
      double ox = (c * r.o[0]) - (s * r.o[2]), oz = (s * r.o[0]) + (c * r.o[2]);
      double dx = (c * r.d[0]) - (s * r.d[2]), dz = (s * r.d[0]) + (c * r.d[2]);
      r.o[0] = ox;
      r.o[2] = oz;
      r.d[0] = dx;
      r.d[2] = dz;
    }
  }
  return r;
}

RT_HD bool sphere_hit_exact(const PrimExact &p, const RayD &r, double tmin, double tmax, bool closed_max,
                            double &t_out, int &front) {
  double center[3] = {p.a[0] + r.time * p.b[0], p.a[1] + r.time * p.b[1], p.a[2] + r.time * p.b[2]};
  double oc[3] = {center[0] - r.o[0], center[1] - r.o[1], center[2] - r.o[2]};
  double a = dot3(r.d, r.d);
  double h = dot3(r.d, oc);
  double c = dot3(oc, oc) - p.s * p.s;
  double disc = h * h - a * c;
  if (disc < 0)
    return false;
  double sqrtd = sqrt(disc);
  double root = (h - sqrtd) / a;
  if (!(tmin < root && (closed_max ? root <= tmax : root < tmax))) {
    root = (h + sqrtd) / a;
    if (!(tmin < root && (closed_max ? root <= tmax : root < tmax)))
      return false;
  }
  t_out = root;
  double hp[3] = {r.o[0] + root * r.d[0], r.o[1] + root * r.d[1], r.o[2] + root * r.d[2]};
  double inv_r = 1 / p.s;
  double outward[3] = {inv_r * (hp[0] - center[0]), inv_r * (hp[1] - center[1]), inv_r * (hp[2] - center[2])};
  front = dot3(r.d, outward) < 0;
  return true;
}

RT_HD bool quad_hit_exact(const PrimExact &p, const RayD &r, double tmin, double tmax, double &t_out, int &front) {
  double denom = dot3(p.n, r.d);
  if (fabs(denom) < 1e-8)
    return false;
  double t = (p.s - dot3(p.n, r.o)) / denom;
  if (!(tmin <= t && t <= tmax))
    return false;
  double ip[3] = {r.o[0] + t * r.d[0], r.o[1] + t * r.d[1], r.o[2] + t * r.d[2]};
  double hp[3] = {ip[0] - p.a[0], ip[1] - p.a[1], ip[2] - p.a[2]};
  double c1[3] = {hp[1] * p.c[2] - hp[2] * p.c[1], hp[2] * p.c[0] - hp[0] * p.c[2], hp[0] * p.c[1] - hp[1] * p.c[0]};
  double c2[3] = {p.b[1] * hp[2] - p.b[2] * hp[1], p.b[2] * hp[0] - p.b[0] * hp[2], p.b[0] * hp[1] - p.b[1] * hp[0]};
  double alpha = dot3(p.w, c1);
  double beta = dot3(p.w, c2);
  if (!(0 <= alpha && alpha <= 1) || !(0 <= beta && beta <= 1))
    return false;
  t_out = t;
  front = dot3(r.d, p.n) < 0;
  return true;
}

// Closest hit of a medium's boundary list (HittableList::hit inside the wrapper chain).
RT_HD bool boundary_hit_exact(const ExactScene &sc, const PrimExact &m, const RayD &world_ray, double tmin,
                              double tmax, double &t_out) {
  bool any = false;
  double closest = tmax;
  RayD r = world_ray;
  int last_xf = -2;
  for (int k = 0; k < m.count; k++) {
    const PrimExact &b = sc.bprims[m.first + k];
    if (b.xform != last_xf) {
      r = to_object_space(sc, b.xform, world_ray);
      last_xf = b.xform;
    }
    double t;
    int front;
    bool ok = b.type == RT_PT_SPHERE ? sphere_hit_exact(b, r, tmin, closest, false, t, front)
                                     : quad_hit_exact(b, r, tmin, closest, t, front);
    if (ok) {
      any = true;
      closest = t;
    }
  }
  t_out = closest;
  return any;
}

RT_HD bool medium_hit_exact(const ExactScene &sc, const PrimExact &m, const RayD &ray, double tmin, double tmax,
                            const RayKey &key, double &t_out) {
  double t1, t2;
  const double inf = (double)RT_INF_F;
  if (!boundary_hit_exact(sc, m, ray, -inf, inf, t1))
    return false;
  if (!boundary_hit_exact(sc, m, ray, t1 + 0.0001, inf, t2))
    return false;
  if (t1 < tmin)
    t1 = tmin;
  if (t2 > tmax)
    t2 = tmax;
  if (t1 >= t2)
    return false;
  if (t1 < 0)
    t1 = 0;
  double ray_length = sqrt(dot3(ray.d, ray.d));
  double distance_inside = (t2 - t1) * ray_length;
  double neg_inv_density = -1.0 / m.s;
  Uniform4 u = philox_uniform4(key.seed, key.pixel, key.sample, key.bounce, (uint32_t)(RT_STREAM_MEDIUM0 + m.medium), 0);
  double hit_distance = neg_inv_density * log((double)u.x);
  if (hit_distance > distance_inside)
    return false;
  t_out = t1 + hit_distance / ray_length;
  return true;
}

// Leaf test with the reference list's tie rule: of two primitives hitting at exactly the same t,
// the one later in the reference's scan order (object, then member order) wins iff it uses the
// closed interval test (quads); see oracle/rt_oracle.c bvh_walk.
// The render path's rule for the primitive a ray starts on (rt_device.h, sphere_hit_from_surface), in FP64: a
// sphere is met at its far root only, and only by a ray that enters it.
RT_HD bool sphere_hit_from_surface_exact(const PrimExact &p, const RayD &r, double tmin, double tmax, bool closed_max,
                                         double &t_out, int &front) {
  double center[3] = {p.a[0] + r.time * p.b[0], p.a[1] + r.time * p.b[1], p.a[2] + r.time * p.b[2]};
  double oc[3] = {center[0] - r.o[0], center[1] - r.o[1], center[2] - r.o[2]};
  double a = dot3(r.d, r.d);
  double h = dot3(r.d, oc);
  if (!(h > 0))
    return false;
  double c = dot3(oc, oc) - p.s * p.s;
  double disc = h * h - a * c;
  double root = (h + sqrt(disc > 0 ? disc : 0)) / a;
  if (!(tmin < root && (closed_max ? root <= tmax : root < tmax)))
    return false;
  t_out = root;
  front = 0; // leaving through the far side
  return true;
}

// start_prim: -1 for the parity hook (the reference's own semantics); the audit of the render path passes the
// primitive the ray starts on.
RT_HD void leaf_test_exact(const ExactScene &sc, int prim, const RayD &ray, double tmin, HitD &hit,
                           const RayKey &key, int start_prim = -1) {
  const PrimExact &p = sc.prims[prim];
  bool have = hit.prim >= 0;
  double t;
  int front = 1;
  bool ok;
  if (p.type == RT_PT_MEDIUM) {
    ok = medium_hit_exact(sc, p, ray, tmin, hit.t, key, t);
    if (ok && have && t == hit.t)
      ok = false;
  } else {
    RayD r = to_object_space(sc, p.xform, ray);
    if (p.type == RT_PT_SPHERE)
      ok = prim == start_prim ? sphere_hit_from_surface_exact(p, r, tmin, hit.t, have, t, front)
                              : sphere_hit_exact(p, r, tmin, hit.t, have, t, front);
    else
      ok = prim != start_prim && quad_hit_exact(p, r, tmin, hit.t, t, front);
    if (ok && have && t == hit.t) {
      const PrimExact &b = sc.prims[hit.prim];
      bool later = p.object > b.object || (p.object == b.object && p.id > b.id);
      bool wins = later ? (p.closed != 0) : (b.closed == 0);
      if (!wins)
        ok = false;
    }
  }
  if (ok) {
    hit.t = t;
    hit.prim = prim;
    hit.front = front;
    hit.object = p.object;
    hit.id = p.id;
  }
}

// start_prim (leaf order, -1 = none): the surface primitive the ray starts on (rt_device.h leaf_test); the parity
// hook passes -1.
RT_HD void traverse_exact(const ExactScene &sc, const RayD &ray, double tmin, double tmax, HitD &hit,
                          const RayKey &key, int start_prim = -1) {
  hit.t = tmax;
  hit.prim = -1;
  hit.front = 0;
  hit.object = -1;
  hit.id = -1;
  double inv[3] = {1.0 / ray.d[0], 1.0 / ray.d[1], 1.0 / ray.d[2]};
  int stack[RT_STACK * 2];
  int sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    int ref = stack[--sp];
    if (ref < 0) {
      leaf_test_exact(sc, ~ref, ray, tmin, hit, key, start_prim);
      continue;
    }
    const float4 *n = sc.nodes + (size_t)ref * RT_NODE_F4;
    float4 lox = n[0], hix = n[1], loy = n[2], hiy = n[3], loz = n[4], hiz = n[5], cr = n[6];
    const float lo[3][4] = {{lox.x, lox.y, lox.z, lox.w}, {loy.x, loy.y, loy.z, loy.w}, {loz.x, loz.y, loz.z, loz.w}};
    const float hi[3][4] = {{hix.x, hix.y, hix.z, hix.w}, {hiy.x, hiy.y, hiy.z, hiy.w}, {hiz.x, hiz.y, hiz.z, hiz.w}};
    const int cref[4] = {f2i(cr.x), f2i(cr.y), f2i(cr.z), f2i(cr.w)};
    for (int c = 0; c < 4; c++) {
      if (cref[c] == RT_EMPTY)
        continue;
      double tn = tmin, tf = hit.t;
      bool ok = true;
      for (int ax = 0; ax < 3; ax++) {
        double t0 = ((double)lo[ax][c] - ray.o[ax]) * inv[ax];
        double t1 = ((double)hi[ax][c] - ray.o[ax]) * inv[ax];
        double a = t0 < t1 ? t0 : t1, b = t0 < t1 ? t1 : t0;
        if (a == a && a > tn) // NaN (0 * inf) leaves the interval unchanged
          tn = a;
        if (b == b && b < tf)
          tf = b;
        if (tf < tn)
          ok = false;
      }
      if (ok && sp < RT_STACK * 2)
        stack[sp++] = cref[c];
    }
  }
}
