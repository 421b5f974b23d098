// rt_flatten.h — host-side flattening of an rt_scene_desc into the device record arrays (pure C++,
// FP64): instance chains baked to world space for the render path, object-space FP64 records for the
// parity path, conservative FP32 leaf boxes for the BVH build.  Included by rt_scene.cu (product) and
// by the host test build in tests/emu.
#pragma once

#include "../../include/rt_b200.h"
#include "rt_bvh.h"
#include "rt_device.h"
#include "rt_exact.h"
#include "rt_bigvec.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

void rt_set_error(const std::string &msg);

namespace rtflat {

// body(first, last) over [0, n) on up to 16 host threads; small ranges run on the caller.  The million-primitive
// scenes spend their creation time in per-primitive FP64 baking, which is independent per primitive.
template <class Body> inline void parallel_for(size_t n, Body body) {
  const size_t kMinPerThread = 16384;
  size_t threads = std::min<size_t>(std::min<size_t>(std::thread::hardware_concurrency(), 16), n / kMinPerThread);
  if (threads <= 1) {
    body((size_t)0, n);
    return;
  }
  std::vector<std::thread> pool;
  size_t chunk = (n + threads - 1) / threads;
  for (size_t t = 0; t < threads; t++) {
    size_t a = t * chunk, b = std::min(n, a + chunk);
    if (a < b)
      pool.emplace_back([=, &body]() { body(a, b); });
  }
  for (std::thread &t : pool)
    t.join();
}

struct D3 {
  double x, y, z;
};
inline D3 d3(const double *p) { return {p[0], p[1], p[2]}; }
inline D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 operator*(double t, D3 a) { return {t * a.x, t * a.y, t * a.z}; }
inline double dotd(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline D3 crossd(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline D3 unitd(D3 a) { // Vec3::normalize (utils/math/Vec3.hpp:141-149)
  double len = std::sqrt(dotd(a, a));
  if (len > 1e-8)
    return (1.0 / len) * a;
  return {1.0, 0.0, 0.0};
}
inline float4 f4(D3 v, float w) { return make_float4((float)v.x, (float)v.y, (float)v.z, w); }

struct BoxD {
  double lo[3], hi[3];
  BoxD() {
    for (int a = 0; a < 3; a++) {
      lo[a] = std::numeric_limits<double>::infinity();
      hi[a] = -std::numeric_limits<double>::infinity();
    }
  }
  void grow(D3 p) {
    const double v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], v[a]);
      hi[a] = std::max(hi[a], v[a]);
    }
  }
  void grow(const BoxD &b) {
    for (int a = 0; a < 3; a++) {
      lo[a] = std::min(lo[a], b.lo[a]);
      hi[a] = std::max(hi[a], b.hi[a]);
    }
  }
};

// FP64 box -> FP32 box that contains it, at least 1e-4 thick on every axis (the reference pads flat
// boxes the same way, optimization/AABB.cpp:167-176) plus two ulps of slack for the FP32 slab test.
inline BuildBox to_build_box(const BoxD &b) {
  BuildBox r;
  for (int a = 0; a < 3; a++) {
    double lo = b.lo[a], hi = b.hi[a];
    if (hi - lo < 1e-4) {
      lo -= 5e-5;
      hi += 5e-5;
    }
    float l = (float)lo, h = (float)hi;
    if ((double)l > lo)
      l = std::nextafterf(l, -INFINITY);
    if ((double)h < hi)
      h = std::nextafterf(h, INFINITY);
    for (int k = 0; k < 2; k++) {
      l = std::nextafterf(l, -INFINITY);
      h = std::nextafterf(h, INFINITY);
    }
    r.lo[a] = l;
    r.hi[a] = h;
  }
  return r;
}

struct Baker {
  const rt_scene_desc *d;
  // object space -> world space through chain xf (innermost wrapper first): RotateY::hit's and
  // Translate::hit's point transforms (RotateY.cpp:64-67, Translate.cpp:25)
  D3 point(int xf, D3 p) const {
    if (xf < 0)
      return p;
    const rt_xform &x = d->xforms[xf];
    for (int k = x.n_ops - 1; k >= 0; k--) {
      const rt_xform_op &op = d->xform_ops[x.first_op + k];
      if (op.type == RT_XF_TRANSLATE)
        p = p + d3(op.offset);
      else
        p = {op.cos_theta * p.x + op.sin_theta * p.z, p.y, -op.sin_theta * p.x + op.cos_theta * p.z};
    }
    return p;
  }
  D3 vector(int xf, D3 v) const {
    if (xf < 0)
      return v;
    const rt_xform &x = d->xforms[xf];
    for (int k = x.n_ops - 1; k >= 0; k--) {
      const rt_xform_op &op = d->xform_ops[x.first_op + k];
      if (op.type == RT_XF_ROTATE_Y)
        v = {op.cos_theta * v.x + op.sin_theta * v.z, v.y, -op.sin_theta * v.x + op.cos_theta * v.z};
    }
    return v;
  }
};

inline float ibits(int i) {
  float f;
  std::memcpy(&f, &i, 4);
  return f;
}

struct Flat {
  BigVec<float4> prims, mats; // per surface / per material: written element by element by the baking threads
  std::vector<float4> bprims, lights, perlin_grad;
  std::vector<unsigned char> perlin_perm;
  std::vector<uint32_t> texels; // image textures, 0x00BBGGRR per texel
  BigVec<PrimExact> ex_prims;
  std::vector<PrimExact> ex_bprims;
  std::vector<XformOpExact> ops;
  std::vector<int> chain_first, chain_count;
  BigVec<BoxD> boxes;
};

// Spheres carry a copy of their material: the header in [2] and, when one colour is all the material needs
// (metal, or a solid texture), that colour in [2].w, [1].w, [3].x - shading a sphere hit then takes one dependent
// fetch (the primitive record) instead of three (record -> header -> colour).  `rec` = the 4 float4 of a record.
template <class Mats> inline void embed_sphere_material(float4 *rec, const Mats &mats) {
  uint32_t typemat = (uint32_t)f2i(rec[3].y);
  size_t m = (size_t)(typemat & 0x0fffffffu) * RT_MAT_F4;
  if ((typemat >> 28) != RT_PT_SPHERE || m + 1 >= mats.size())
    return;
  float4 m0 = mats[m], a = mats[m + 1];
  if (f2i(m0.x) == RT_MAT_METAL || f2i(m0.y) == RT_DTEX_SOLID) {
    m0.w = a.x;
    rec[1].w = a.y;
    rec[3].x = a.z;
  }
  rec[2] = m0;
}

// rec: the 4 float4 of the render-path record, exact: the FP64 parity record, box: grown by the primitive's bounds
inline void bake_sphere(const Baker &bk, const rt_sphere &s, int id, int material, float4 *rec, PrimExact &exact, BoxD &box) {
  D3 c0 = bk.point(s.xform, d3(s.center0));
  D3 dir = bk.vector(s.xform, d3(s.center_dir));
  double r = std::fmax(0.0, s.radius);
  int typemat = (RT_PT_SPHERE << 28) | (material < 0 ? 0 : material);
  rec[0] = f4(c0, (float)r);
  rec[1] = f4(dir, 0.f);
  rec[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  rec[3] = make_float4(0.f, ibits(typemat), ibits(id), ibits(s.object));
  PrimExact e{};
  std::memcpy(e.a, s.center0, sizeof e.a);
  std::memcpy(e.b, s.center_dir, sizeof e.b);
  e.s = r;
  e.type = RT_PT_SPHERE;
  e.xform = s.xform;
  e.id = id;
  e.object = s.object;
  e.closed = 0;
  e.medium = -1;
  exact = e;
  D3 c1 = c0 + dir;
  D3 rv = {r, r, r};
  box.grow(c0 - rv);
  box.grow(c0 + rv);
  box.grow(c1 - rv);
  box.grow(c1 + rv);
}

inline void push_sphere(const Baker &bk, const rt_sphere &s, int id, int material, std::vector<float4> &fast,
                        std::vector<PrimExact> &exact, BoxD &box) {
  fast.resize(fast.size() + RT_PRIM_F4);
  exact.emplace_back();
  bake_sphere(bk, s, id, material, &fast[fast.size() - RT_PRIM_F4], exact.back(), box);
}

inline void bake_quad(const Baker &bk, const rt_quad &q, int id, int material, float4 *rec, PrimExact &exact, BoxD &box) {
  // world-space record for the render path
  D3 Q = bk.point(q.xform, d3(q.corner));
  D3 u = bk.vector(q.xform, d3(q.u)), v = bk.vector(q.xform, d3(q.v));
  D3 n = crossd(u, v);
  D3 normal = unitd(n);
  double D = dotd(normal, Q);
  D3 w = (1.0 / dotd(n, n)) * n;
  D3 A = crossd(v, w), B = crossd(w, u);
  int typemat = (RT_PT_QUAD << 28) | (material < 0 ? 0 : material);
  rec[0] = f4(normal, (float)D);
  rec[1] = f4(A, (float)Q.x);
  rec[2] = f4(B, (float)Q.y);
  rec[3] = make_float4((float)Q.z, ibits(typemat), ibits(id), ibits(q.object));
  // object-space record for the FP64 parity path: Plane's constructor (Plane.cpp:6-21)
  PrimExact e{};
  D3 qo = d3(q.corner), uo = d3(q.u), vo = d3(q.v);
  D3 no = crossd(uo, vo);
  D3 normal_o = unitd(no);
  D3 wo = (1 / dotd(no, no)) * no;
  std::memcpy(e.a, q.corner, sizeof e.a);
  std::memcpy(e.b, q.u, sizeof e.b);
  std::memcpy(e.c, q.v, sizeof e.c);
  e.n[0] = normal_o.x, e.n[1] = normal_o.y, e.n[2] = normal_o.z;
  e.w[0] = wo.x, e.w[1] = wo.y, e.w[2] = wo.z;
  e.s = dotd(normal_o, qo);
  e.type = RT_PT_QUAD;
  e.xform = q.xform;
  e.id = id;
  e.object = q.object;
  e.closed = 1;
  e.medium = -1;
  exact = e;
  box.grow(Q);
  box.grow(Q + u);
  box.grow(Q + v);
  box.grow(Q + u + v);
}

inline void push_quad(const Baker &bk, const rt_quad &q, int id, int material, std::vector<float4> &fast,
                      std::vector<PrimExact> &exact, BoxD &box) {
  fast.resize(fast.size() + RT_PRIM_F4);
  exact.emplace_back();
  bake_quad(bk, q, id, material, &fast[fast.size() - RT_PRIM_F4], exact.back(), box);
}

constexpr double kMaxCoordinate = 1e18; // FP32 box areas stay finite: 6 * (2e18)^2 < FLT_MAX

inline int fail_invalid(const std::string &msg) {
  rt_set_error("invalid scene: " + msg);
  return RT_ERR_INVALID;
}

inline int flatten(const rt_scene_desc *d, Flat &f) {
  if (!d)
    return fail_invalid("null description");
  if (d->n_spheres < 0 || d->n_quads < 0 || d->n_media < 0 || d->n_materials < 0 || d->n_textures < 0 ||
      d->n_perlins < 0 || d->n_lights < 0 || d->n_xforms < 0 || d->n_xform_ops < 0)
    return fail_invalid("negative count");
  if (d->n_materials >= (1 << 28))
    return fail_invalid("too many materials");
  Baker bk{d};
  auto check_xf = [&](int xf) { return xf >= -1 && xf < d->n_xforms; };
  auto check_mat = [&](int m) { return m >= 0 && m < d->n_materials; };
  for (int i = 0; i < d->n_xforms; i++) {
    const rt_xform &x = d->xforms[i];
    if (x.first_op < 0 || x.n_ops < 0 || x.first_op + x.n_ops > d->n_xform_ops)
      return fail_invalid("instance chain out of range");
    f.chain_first.push_back(x.first_op);
    f.chain_count.push_back(x.n_ops);
  }
  for (int i = 0; i < d->n_xform_ops; i++) {
    const rt_xform_op &op = d->xform_ops[i];
    if (op.type != RT_XF_TRANSLATE && op.type != RT_XF_ROTATE_Y)
      return fail_invalid("unknown instance op");
    XformOpExact e{};
    e.type = op.type;
    std::memcpy(e.offset, op.offset, sizeof e.offset);
    e.sin_theta = op.sin_theta;
    e.cos_theta = op.cos_theta;
    f.ops.push_back(e);
  }

  // surfaces: indices are checked and record slots handed out in one cheap serial pass, the FP64 baking of the
  // records (instance chain, parity record, bounds) then runs over all host threads
  std::vector<int> sphere_slot(d->n_spheres, -1), quad_slot(d->n_quads, -1);
  int n_surface = 0;
  for (int i = 0; i < d->n_spheres; i++) {
    const rt_sphere &s = d->spheres[i];
    if (!check_xf(s.xform))
      return fail_invalid("sphere instance chain index");
    if (s.flags & RT_PRIM_BOUNDARY)
      continue;
    if (!check_mat(s.material))
      return fail_invalid("sphere material index");
    sphere_slot[i] = n_surface++;
  }
  for (int i = 0; i < d->n_quads; i++) {
    const rt_quad &q = d->quads[i];
    if (!check_xf(q.xform))
      return fail_invalid("quad instance chain index");
    if (q.flags & RT_PRIM_BOUNDARY)
      continue;
    if (!check_mat(q.material))
      return fail_invalid("quad material index");
    quad_slot[i] = n_surface++;
  }
  f.prims.resize((size_t)n_surface * RT_PRIM_F4);
  f.ex_prims.resize(n_surface);
  f.boxes.resize(n_surface);
  parallel_for((size_t)d->n_spheres, [&](size_t a, size_t b) {
    for (size_t i = a; i < b; i++)
      if (sphere_slot[i] >= 0) {
        const int k = sphere_slot[i];
        f.boxes[k] = BoxD();
        bake_sphere(bk, d->spheres[i], (int)i, d->spheres[i].material, &f.prims[(size_t)k * RT_PRIM_F4], f.ex_prims[k], f.boxes[k]);
      }
  });
  parallel_for((size_t)d->n_quads, [&](size_t a, size_t b) {
    for (size_t i = a; i < b; i++)
      if (quad_slot[i] >= 0) {
        const int k = quad_slot[i];
        f.boxes[k] = BoxD();
        bake_quad(bk, d->quads[i], d->n_spheres + (int)i, d->quads[i].material, &f.prims[(size_t)k * RT_PRIM_F4], f.ex_prims[k],
                  f.boxes[k]);
      }
  });
  // constant media: one leaf each, boundary records on the side
  for (int m = 0; m < d->n_media; m++) {
    const rt_medium &md = d->media[m];
    if (!check_mat(md.material))
      return fail_invalid("medium phase-function material index");
    if (!(md.density > 0))
      return fail_invalid("medium density must be positive");
    int limit = md.shape == RT_SHAPE_SPHERE ? d->n_spheres : d->n_quads;
    if ((md.shape != RT_SHAPE_SPHERE && md.shape != RT_SHAPE_QUAD) || md.first_prim < 0 || md.n_prims <= 0 ||
        md.first_prim + md.n_prims > limit)
      return fail_invalid("medium boundary range");
    int first = (int)f.ex_bprims.size();
    BoxD box;
    for (int k = 0; k < md.n_prims; k++) {
      if (md.shape == RT_SHAPE_SPHERE)
        push_sphere(bk, d->spheres[md.first_prim + k], md.first_prim + k, -1, f.bprims, f.ex_bprims, box);
      else
        push_quad(bk, d->quads[md.first_prim + k], d->n_spheres + md.first_prim + k, -1, f.bprims, f.ex_bprims, box);
    }
    int id = d->n_spheres + d->n_quads + m;
    int typemat = (RT_PT_MEDIUM << 28) | md.material;
    f.prims.push_back(make_float4((float)(-1.0 / md.density), ibits(first), ibits(md.n_prims), ibits(m)));
    f.prims.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
    f.prims.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
    f.prims.push_back(make_float4(0.f, ibits(typemat), ibits(id), ibits(md.object)));
    PrimExact e{};
    e.s = md.density;
    e.type = RT_PT_MEDIUM;
    e.xform = -1;
    e.id = id;
    e.object = md.object;
    e.closed = 0;
    e.medium = m;
    e.first = first;
    e.count = md.n_prims;
    f.ex_prims.push_back(e);
    f.boxes.push_back(box);
  }

  // image textures: every image repacked to one 32-bit word per texel, all images back to back
  std::vector<size_t> image_offset;
  if (d->n_images < 0)
    return fail_invalid("negative image count");
  for (int i = 0; i < d->n_images; i++) {
    const rt_image &im = d->images[i];
    if (im.width < 0 || im.height < 0 || ((int64_t)im.width * im.height > 0 && !im.rgb) ||
        (int64_t)im.width * im.height > ((int64_t)1 << 28))
      return fail_invalid("image texture size / data");
    image_offset.push_back(f.texels.size());
    for (int64_t k = 0; k < (int64_t)im.width * im.height; k++)
      f.texels.push_back((uint32_t)im.rgb[3 * k] | ((uint32_t)im.rgb[3 * k + 1] << 8) | ((uint32_t)im.rgb[3 * k + 2] << 16));
  }
  if (f.texels.size() > ((size_t)1 << 31))
    return fail_invalid("image textures exceed 2^31 texels");

  // materials with their texture folded in (a million-sphere scene has a million of them: all host threads)
  f.mats.resize((size_t)d->n_materials * RT_MAT_F4);
  auto bake_material = [&](int i) -> int {
    const rt_material &m = d->materials[i];
    float4 m0 = make_float4(ibits(m.type), ibits(RT_DTEX_SOLID), 0.f, 0.f);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (m.type == RT_MAT_METAL) {
      m0.z = (float)m.fuzz;
      a = f4(d3(m.albedo), 0.f);
    } else if (m.type == RT_MAT_DIELECTRIC) {
      m0.z = (float)m.ior;
    } else if (m.type == RT_MAT_LAMBERTIAN || m.type == RT_MAT_DIFFUSE_LIGHT || m.type == RT_MAT_ISOTROPIC) {
      if (m.texture < 0 || m.texture >= d->n_textures)
        return fail_invalid("material texture index");
      const rt_texture &t = d->textures[m.texture];
      if (t.type == RT_TEX_SOLID) {
        a = f4(d3(t.color), 0.f);
      } else if (t.type == RT_TEX_CHECKER) {
        if (t.even < 0 || t.even >= d->n_textures || t.odd < 0 || t.odd >= d->n_textures)
          return fail_invalid("checker child texture index");
        if (d->textures[t.even].type != RT_TEX_SOLID || d->textures[t.odd].type != RT_TEX_SOLID) {
          rt_set_error("checker textures whose children are not solid colours are not supported");
          return RT_ERR_UNSUPPORTED;
        }
        m0.y = ibits(RT_DTEX_CHECKER);
        m0.z = (float)t.scale;
        a = f4(d3(d->textures[t.even].color), 0.f);
        b = f4(d3(d->textures[t.odd].color), 0.f);
      } else if (t.type == RT_TEX_NOISE) {
        if (t.perlin < 0 || t.perlin >= d->n_perlins)
          return fail_invalid("noise texture perlin index");
        m0.y = ibits(RT_DTEX_NOISE);
        m0.z = (float)t.scale;
        m0.w = ibits(t.perlin);
      } else if (t.type == RT_TEX_IMAGE) {
        if (t.perlin < 0 || t.perlin >= d->n_images)
          return fail_invalid("image texture index");
        m0.y = ibits(RT_DTEX_IMAGE);
        m0.z = ibits(d->images[t.perlin].width);
        m0.w = ibits(d->images[t.perlin].height);
        a = make_float4(ibits((int)image_offset[t.perlin]), 0.f, 0.f, 0.f);
      } else {
        return fail_invalid("unknown texture type");
      }
    } else {
      return fail_invalid("unknown material type");
    }
    f.mats[(size_t)i * RT_MAT_F4 + 0] = m0;
    f.mats[(size_t)i * RT_MAT_F4 + 1] = a;
    f.mats[(size_t)i * RT_MAT_F4 + 2] = b;
    return RT_OK;
  };
  {
    int bad = -1;
    std::mutex guard;
    parallel_for((size_t)d->n_materials, [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; i++)
        if (bake_material((int)i) != RT_OK) { // the message a worker thread sets is its own (thread-local)
          std::lock_guard<std::mutex> lock(guard);
          if (bad < 0 || (int)i < bad)
            bad = (int)i;
        }
    });
    if (bad >= 0) { // re-run the first failing material on this thread for its status and message
      int st = bake_material(bad);
      if (st != RT_OK)
        return st;
    }
  }

  parallel_for(f.prims.size() / RT_PRIM_F4, [&](size_t a, size_t b) {
    for (size_t r = a; r < b; r++)
      embed_sphere_material(&f.prims[r * RT_PRIM_F4], f.mats);
  });

  // Geometry the FP32 build cannot order is refused here, not rendered: a NaN / infinite coordinate, or a box
  // whose FP32 surface area overflows (|coordinate| above ~1e18), would leave the SAH sweep without a finite
  // cost and the Morton codes without a scale.
  {
    std::atomic<bool> bad_box{false};
    parallel_for(f.boxes.size(), [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; i++)
        for (int a = 0; a < 3; a++)
          if (!(std::fabs(f.boxes[i].lo[a]) <= kMaxCoordinate) || !(std::fabs(f.boxes[i].hi[a]) <= kMaxCoordinate))
            bad_box.store(true, std::memory_order_relaxed);
    });
    if (bad_box.load())
      return fail_invalid("primitive with a non-finite or too large coordinate (|x| must stay below 1e18)");
  }

  for (int i = 0; i < d->n_perlins; i++) {
    const rt_perlin &p = d->perlins[i];
    for (int k = 0; k < RT_PERLIN_POINTS; k++)
      f.perlin_grad.push_back(f4(d3(p.rand_vec[k]), 0.f));
    for (const int32_t *perm : {p.perm_x, p.perm_y, p.perm_z})
      for (int k = 0; k < RT_PERLIN_POINTS; k++) {
        if (perm[k] < 0 || perm[k] > 255)
          return fail_invalid("perlin permutation entry");
        f.perlin_perm.push_back((unsigned char)perm[k]);
      }
  }

  for (int i = 0; i < d->n_lights; i++) {
    const rt_light &l = d->lights[i];
    if (l.xform != -1) {
      rt_set_error("instanced light proxies are not supported");
      return RT_ERR_UNSUPPORTED;
    }
    if (l.shape == RT_SHAPE_QUAD) {
      D3 Q = d3(l.a), u = d3(l.b), v = d3(l.c);
      D3 n = crossd(u, v);
      D3 normal = unitd(n);
      D3 w = (1.0 / dotd(n, n)) * n;
      f.lights.push_back(f4(Q, ibits(1)));
      f.lights.push_back(f4(u, (float)std::sqrt(dotd(n, n))));
      f.lights.push_back(f4(v, 0.f));
      f.lights.push_back(f4(normal, (float)dotd(normal, Q)));
      f.lights.push_back(f4(crossd(v, w), 0.f));
      f.lights.push_back(f4(crossd(w, u), 0.f));
    } else if (l.shape == RT_SHAPE_SPHERE) {
      f.lights.push_back(f4(d3(l.a), ibits(0)));
      f.lights.push_back(make_float4((float)std::fmax(0.0, l.radius), 0.f, 0.f, 0.f));
      for (int k = 0; k < 4; k++)
        f.lights.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
    } else {
      return fail_invalid("unknown light shape");
    }
  }
  return RT_OK;
}


} // namespace rtflat
