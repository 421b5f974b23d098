// oracle/ref_harness.cpp — TEST INFRASTRUCTURE, not product code.
//
// A driver that is linked against the UNMODIFIED reference sources where they lie under
// /root/reference/src (see oracle/Makefile; nothing is copied into this repository) and exposes the
// reference's CPU path through a small C interface for ctypes:
//   * fixture scenes built from reference classes (Sphere, Plane, make_box, RotateY, Translate,
//     ConstantMedium, materials, textures) — while every object is created, the same parameters are
//     recorded into the flat rt_scene_desc arrays of include/rt_b200.h, so the CUDA library and the C
//     restatement (oracle/rt_oracle.c) consume exactly the data the reference renders;
//   * closest-hit queries through Hittable::hit with a tagging decorator that reports the index of
//     the top-level world object (the reference has no primitive ids, HitRecord.hpp:14-44);
//   * Camera::get_ray / Camera::ray_color (reached by subclassing Camera, Camera.hpp:132,156) for
//     deterministic single-thread renders (random_engine().seed(k), Utility.hpp:16-19) and for the
//     multi-threaded CPU baseline.
//
// Harness-level deviations from the reference's shipped behaviour (all documented in DESIGN.md):
//   * DielectricMaterial::scatter leaves Ray::m_time uninitialised (DielectricMaterial.cpp:82,
//     Ray.cpp:6-7).  Dielectrics are wrapped in TimeFixMaterial, which calls the reference scatter and
//     then re-attaches the incoming ray's time (what the reference's CUDA path does, Material.cuh:139).
//   * Sphere::pdf_value has the same uninitialised-time read (Sphere.cpp:149); sphere light proxies use
//     FixedTimeSphereLight below (same formula, probe time 0).
//   * The reference's flattened BVH is undefined behaviour above 1000 nodes (BVHNode.cpp:327,339-360),
//     so for more than 400 objects the tree is assembled from reference BVHNode(left,right) nodes by
//     median split (BVHNode.hpp:92-100); traversal, AABB::hit and primitive tests stay reference code.
//   * The parallel driver hands out scanlines to std::threads with an atomic counter instead of the
//     reference ThreadPool (one job per pixel overflows its 1024-slot deque for width > 1024,
//     StaticCamera.cpp:68-90).
#include "core/Hittable.hpp"
#include "core/HittableList.hpp"
#include "core/HitRecord.hpp"
#include "core/Ray.hpp"
#include "core/ScatterRecord.hpp"
#include "core/camera/Camera.hpp"
#include "optimization/AABB.hpp"
#include "optimization/BVHNode.hpp"
#include "scene/materials/DielectricMaterial.hpp"
#include "scene/materials/DiffuseLightMaterial.hpp"
#include "scene/materials/IsotropicMaterial.hpp"
#include "scene/materials/LambertianMaterial.hpp"
#include "scene/materials/MetalMaterial.hpp"
#include "scene/mediums/ConstantMedium.hpp"
#include "scene/objects/Plane.hpp"
#include "scene/objects/PlaneUtility.hpp"
#include "scene/objects/RotateY.hpp"
#include "scene/objects/Sphere.hpp"
#include "scene/objects/Translate.hpp"
#include "scene/textures/CheckerTexture.hpp"
#include "scene/textures/NoiseTexture.hpp"
#include "scene/textures/SolidColorTexture.hpp"
#include "utils/ColorUtility.hpp"
#include "utils/math/Vec3Utility.hpp"

#include "../include/rt_b200.h"

#ifdef USE_CUDA // the GPU comparator build (oracle/_ref_gpu): the reference's own CUDA path
#include "core/camera/CameraKernelWrappers.cuh"
#include "scene/CudaSceneInitialization.cuh"
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local int tl_last_object = -1;
thread_local uint64_t tl_segments = 0;

// Reports which top-level object produced the last successful hit.  The last success on a
// shrinking interval is the closest hit (HittableList.cpp:32-38, BVHNode.cpp:420-425).
class Tagged : public Hittable {
public:
  Tagged(HittablePtr inner, int id) : m_inner(std::move(inner)), m_id(id) {}
  AABB get_bounding_box() const override { return m_inner->get_bounding_box(); }
  bool hit(const Ray &ray, Interval t_values, HitRecord &record) const override {
    if (!m_inner->hit(ray, t_values, record))
      return false;
    tl_last_object = m_id;
    return true;
  }
  double pdf_value(const Point3 &o, const Vec3 &d) const override { return m_inner->pdf_value(o, d); }
  Vec3 random(const Point3 &o) const override { return m_inner->random(o); }
  std::string json() const override { return "{}"; }

private:
  HittablePtr m_inner;
  int m_id;
};

// Counts ray segments: Camera::ray_color calls world.hit exactly once per segment (Camera.cpp:242).
class SegmentCounter : public Hittable {
public:
  explicit SegmentCounter(HittablePtr inner) : m_inner(std::move(inner)) {}
  AABB get_bounding_box() const override { return m_inner->get_bounding_box(); }
  bool hit(const Ray &ray, Interval t_values, HitRecord &record) const override {
    ++tl_segments;
    return m_inner->hit(ray, t_values, record);
  }
  std::string json() const override { return "{}"; }

private:
  HittablePtr m_inner;
};

// See header comment: reference scatter + incoming ray time re-attached.
class TimeFixMaterial : public Material {
public:
  explicit TimeFixMaterial(MaterialPtr inner) : m_inner(std::move(inner)) {}
  Color emitted(const Ray &r, const HitRecord &rec, double u, double v, const Point3 &p) const override {
    return m_inner->emitted(r, rec, u, v, p);
  }
  bool scatter(const Ray &r, const HitRecord &rec, ScatterRecord &s) const override {
    bool ok = m_inner->scatter(r, rec, s);
    if (ok && s.skip_pdf)
      s.skip_pdf_ray = Ray(s.skip_pdf_ray.origin(), s.skip_pdf_ray.direction(), r.time());
    return ok;
  }
  double scattering_pdf(const Ray &r, const HitRecord &rec, const Ray &s) const override {
    return m_inner->scattering_pdf(r, rec, s);
  }
  std::string json() const override { return m_inner->json(); }

private:
  MaterialPtr m_inner;
};

// Sphere::pdf_value builds its probe ray with the 2-argument Ray constructor, which leaves m_time
// uninitialised (Sphere.cpp:149, Ray.cpp:6-7); the value it then reads is whatever the stack held, so
// the reference's own renders differ between code paths (list vs BVH world) whenever that garbage is
// a NaN.  Sphere light proxies therefore use this subclass: the reference formula (Sphere.cpp:145-157)
// with the probe ray's time pinned to 0 (light proxies are static, so any finite time is equivalent).
class FixedTimeSphereLight : public Sphere {
public:
  using Sphere::Sphere;
  double pdf_value(const Point3 &origin, const Vec3 &direction) const override {
    HitRecord record;
    if (!this->hit(Ray(origin, direction, 0.0), Interval(0.001, INF), record))
      return 0;
    double radius = get_radius();
    double dist_squared = (get_center().at(0) - origin).length_squared();
    double cos_theta_max = std::sqrt(1 - radius * radius / dist_squared);
    double solid_angle = 2 * PI * (1 - cos_theta_max);
    return 1 / solid_angle;
  }
};

void set3(double *dst, const Vec3 &v) {
  dst[0] = v.x();
  dst[1] = v.y();
  dst[2] = v.z();
}

struct Mat {
  MaterialPtr ptr;
  int idx;
};
struct Tex {
  TexturePtr ptr;
  int idx;
};

// Builds reference objects and records the same parameters into flat arrays.
struct RefScene {
  HittableList world;   // tagged top-level objects, in main.cpp order
  HittableList lights;
  CameraConfig cam;
  std::vector<rt_sphere> spheres;
  std::vector<rt_quad> quads;
  std::vector<rt_xform_op> xform_ops;
  std::vector<rt_xform> xforms;
  std::vector<rt_medium> media;
  std::vector<rt_material> materials;
  std::vector<rt_texture> textures;
  std::vector<rt_perlin> perlins;
  std::vector<rt_light> light_recs;
  rt_scene_desc desc;
  int n_objects = 0;
  // render-time acceleration structure (built lazily)
  std::shared_ptr<HittableList> bvh_world;
  // plain = true: only reference classes in the object graph (no tagging / time-fix decorators), as the
  // reference's own scene functions build it - what its CPU -> CUDA converters understand (ref_gpu_frame)
  bool plain = false;

  // ---- textures ----
  Tex solid(const Color &c) {
    rt_texture t{};
    t.type = RT_TEX_SOLID;
    t.even = t.odd = t.perlin = -1;
    set3(t.color, c);
    textures.push_back(t);
    return {std::make_shared<SolidColorTexture>(c), int(textures.size()) - 1};
  }
  Tex checker(double scale, const Color &c1, const Color &c2) {
    Tex e = solid(c1), o = solid(c2);
    rt_texture t{};
    t.type = RT_TEX_CHECKER;
    t.even = e.idx;
    t.odd = o.idx;
    t.perlin = -1;
    t.scale = scale;
    textures.push_back(t);
    return {std::make_shared<CheckerTexture>(scale, e.ptr, o.ptr), int(textures.size()) - 1};
  }
  int add_perlin(const PerlinNoise &p) {
    rt_perlin r{};
    for (int i = 0; i < RT_PERLIN_POINTS; i++) {
      set3(r.rand_vec[i], p.rand_vec()[i]);
      r.perm_x[i] = p.perm_x()[i];
      r.perm_y[i] = p.perm_y()[i];
      r.perm_z[i] = p.perm_z()[i];
    }
    perlins.push_back(r);
    return int(perlins.size()) - 1;
  }
  Tex noise(double scale, PerlinNoise &p, int perlin_idx) {
    rt_texture t{};
    t.type = RT_TEX_NOISE;
    t.even = t.odd = -1;
    t.perlin = perlin_idx;
    t.scale = scale;
    textures.push_back(t);
    return {std::make_shared<NoiseTexture>(scale, p), int(textures.size()) - 1};
  }

  // ---- materials ----
  Mat push_mat(MaterialPtr p, int type, int tex, const Color &albedo, double fuzz, double ior) {
    rt_material m{};
    m.type = type;
    m.texture = tex;
    set3(m.albedo, albedo);
    m.fuzz = fuzz;
    m.ior = ior;
    materials.push_back(m);
    return {std::move(p), int(materials.size()) - 1};
  }
  Mat lambertian(const Color &c) {
    Tex t = solid(c);
    return push_mat(std::make_shared<LambertianMaterial>(t.ptr), RT_MAT_LAMBERTIAN, t.idx, Color(), 0, 0);
  }
  Mat lambertian(const Tex &t) {
    return push_mat(std::make_shared<LambertianMaterial>(t.ptr), RT_MAT_LAMBERTIAN, t.idx, Color(), 0, 0);
  }
  Mat metal(const Color &c, double fuzz) {
    return push_mat(std::make_shared<MetalMaterial>(c, fuzz), RT_MAT_METAL, -1, c, fuzz, 0);
  }
  Mat dielectric(double ior) {
    MaterialPtr d = std::make_shared<DielectricMaterial>(ior);
    return push_mat(plain ? d : std::make_shared<TimeFixMaterial>(d), RT_MAT_DIELECTRIC, -1, Color(), 0, ior);
  }
  Mat diffuse_light(const Color &c) {
    Tex t = solid(c);
    return push_mat(std::make_shared<DiffuseLightMaterial>(t.ptr), RT_MAT_DIFFUSE_LIGHT, t.idx, Color(), 0, 0);
  }
  Mat isotropic(const Color &c) {
    Tex t = solid(c);
    return push_mat(std::make_shared<IsotropicMaterial>(t.ptr), RT_MAT_ISOTROPIC, t.idx, Color(), 0, 0);
  }

  // ---- transforms: chain given outermost first, e.g. {translate(off), rotate_y(a)} ----
  struct XfSpec {
    int type;
    Vec3 offset;
    double angle;
  };
  int add_xform(const std::vector<XfSpec> &chain) {
    if (chain.empty())
      return -1;
    rt_xform x{int(xform_ops.size()), int(chain.size())};
    for (const XfSpec &s : chain) {
      rt_xform_op op{};
      op.type = s.type;
      if (s.type == RT_XF_TRANSLATE)
        set3(op.offset, s.offset);
      else {
        op.angle_deg = s.angle;
        double radians = degrees_to_radians(s.angle); // RotateY.cpp:6-8
        op.sin_theta = std::sin(radians);
        op.cos_theta = std::cos(radians);
      }
      xform_ops.push_back(op);
    }
    xforms.push_back(x);
    return int(xforms.size()) - 1;
  }
  static HittablePtr wrap(HittablePtr obj, const std::vector<XfSpec> &chain) {
    for (int i = int(chain.size()) - 1; i >= 0; --i) { // innermost wrapper first
      if (chain[i].type == RT_XF_TRANSLATE)
        obj = std::make_shared<Translate>(obj, chain[i].offset);
      else
        obj = std::make_shared<RotateY>(obj, chain[i].angle);
    }
    return obj;
  }

  // ---- primitives (record only; return the reference object) ----
  HittablePtr make_sphere(const Point3 &c0, const Point3 &c1, bool moving, double r, const Mat &m, int xf,
                          int object, int flags) {
    rt_sphere s{};
    set3(s.center0, c0);
    Vec3 dir = moving ? (c1 - c0) : Vec3(0, 0, 0);
    set3(s.center_dir, dir);
    s.radius = std::fmax(0, r);
    s.material = m.idx;
    s.xform = xf;
    s.object = object;
    s.flags = flags;
    spheres.push_back(s);
    if (moving)
      return std::make_shared<Sphere>(c0, c1, r, m.ptr);
    return std::make_shared<Sphere>(c0, r, m.ptr);
  }
  HittablePtr make_quad(const Point3 &q, const Vec3 &u, const Vec3 &v, const Mat &m, int xf, int object,
                        int flags) {
    rt_quad r{};
    set3(r.corner, q);
    set3(r.u, u);
    set3(r.v, v);
    r.material = m.idx;
    r.xform = xf;
    r.object = object;
    r.flags = flags;
    quads.push_back(r);
    return std::make_shared<Plane>(q, u, v, m.ptr);
  }
  // make_box (PlaneUtility.hpp:11-40): records the six quads in the reference's order.
  HittablePtr make_box_rec(const Point3 &a, const Point3 &b, const Mat &m, int xf, int object, int flags) {
    std::shared_ptr<HittableList> box = make_box(a, b, m.ptr);
    for (const HittablePtr &side : box->get_objects()) {
      const Plane *p = dynamic_cast<const Plane *>(side.get());
      rt_quad r{};
      set3(r.corner, p->get_corner());
      set3(r.u, p->get_u_side());
      set3(r.v, p->get_v_side());
      r.material = m.idx;
      r.xform = xf;
      r.object = object;
      r.flags = flags;
      quads.push_back(r);
    }
    return box;
  }

  // ---- top-level world objects ----
  void add_object(HittablePtr obj) {
    if (plain)
      world.add(std::move(obj));
    else
      world.add(std::make_shared<Tagged>(std::move(obj), n_objects));
    ++n_objects;
  }
  void add_sphere(const Point3 &c, double r, const Mat &m) {
    add_object(make_sphere(c, c, false, r, m, -1, n_objects, 0));
  }
  void add_moving_sphere(const Point3 &c0, const Point3 &c1, double r, const Mat &m) {
    add_object(make_sphere(c0, c1, true, r, m, -1, n_objects, 0));
  }
  void add_quad(const Point3 &q, const Vec3 &u, const Vec3 &v, const Mat &m) {
    add_object(make_quad(q, u, v, m, -1, n_objects, 0));
  }
  void add_box(const Point3 &a, const Point3 &b, const Mat &m, const std::vector<XfSpec> &chain) {
    int xf = add_xform(chain);
    add_object(wrap(make_box_rec(a, b, m, xf, n_objects, 0), chain));
  }
  void add_box_medium(const Point3 &a, const Point3 &b, const std::vector<XfSpec> &chain, double density,
                      const Color &albedo) {
    int xf = add_xform(chain);
    Mat none{MaterialPtr(), -1};
    int first = int(quads.size());
    HittablePtr boundary = wrap(make_box_rec(a, b, none, xf, n_objects, RT_PRIM_BOUNDARY), chain);
    Mat phase = isotropic(albedo);
    rt_medium md{};
    md.density = density;
    md.shape = RT_SHAPE_QUAD;
    md.first_prim = first;
    md.n_prims = 6;
    md.material = phase.idx;
    md.object = n_objects;
    media.push_back(md);
    add_object(std::make_shared<ConstantMedium>(boundary, density, phase.ptr));
  }
  void add_sphere_medium(const Point3 &c, double r, double density, const Color &albedo) {
    Mat none{MaterialPtr(), -1};
    int first = int(spheres.size());
    HittablePtr boundary = make_sphere(c, c, false, r, none, -1, n_objects, RT_PRIM_BOUNDARY);
    Mat phase = isotropic(albedo);
    rt_medium md{};
    md.density = density;
    md.shape = RT_SHAPE_SPHERE;
    md.first_prim = first;
    md.n_prims = 1;
    md.material = phase.idx;
    md.object = n_objects;
    media.push_back(md);
    add_object(std::make_shared<ConstantMedium>(boundary, density, phase.ptr));
  }

  // ---- lights (geometry proxies with a null material, main.cpp:57-61) ----
  void add_light_quad(const Point3 &q, const Vec3 &u, const Vec3 &v) {
    rt_light l{};
    l.shape = RT_SHAPE_QUAD;
    l.xform = -1;
    set3(l.a, q);
    set3(l.b, u);
    set3(l.c, v);
    light_recs.push_back(l);
    lights.add(std::make_shared<Plane>(q, u, v, MaterialPtr()));
  }
  void add_light_sphere(const Point3 &c, double r) {
    rt_light l{};
    l.shape = RT_SHAPE_SPHERE;
    l.xform = -1;
    set3(l.a, c);
    l.radius = r;
    light_recs.push_back(l);
    if (plain)
      lights.add(std::make_shared<Sphere>(c, r, MaterialPtr()));
    else
      lights.add(std::make_shared<FixedTimeSphereLight>(c, r, MaterialPtr()));
  }

  void finalize() {
    std::memset(&desc, 0, sizeof desc);
    desc.n_spheres = int(spheres.size());
    desc.n_quads = int(quads.size());
    desc.n_xform_ops = int(xform_ops.size());
    desc.n_xforms = int(xforms.size());
    desc.n_media = int(media.size());
    desc.n_materials = int(materials.size());
    desc.n_textures = int(textures.size());
    desc.n_perlins = int(perlins.size());
    desc.n_lights = int(light_recs.size());
    desc.n_objects = n_objects;
    desc.spheres = spheres.data();
    desc.quads = quads.data();
    desc.xform_ops = xform_ops.data();
    desc.xforms = xforms.data();
    desc.media = media.data();
    desc.materials = materials.data();
    desc.textures = textures.data();
    desc.perlins = perlins.data();
    desc.lights = light_recs.data();
  }
};

using XF = RefScene::XfSpec;
XF translate(const Vec3 &o) { return {RT_XF_TRANSLATE, o, 0}; }
XF rotate_y(double a) { return {RT_XF_ROTATE_Y, Vec3(), a}; }

// ------------------------------------------------------------------------------------------------
// Fixture scenes.  `spheres` and `cornell` restate main.cpp:21-131 (main.cpp itself cannot be linked:
// it pulls in SDL3 through DynamicCamera.hpp); expressions that draw random numbers keep main.cpp's
// exact form so that the compiler evaluates the draws in the same order.
// ------------------------------------------------------------------------------------------------

// populate_bouncing_spheres_scene (main.cpp:73-131); half = 11 is the reference's scene.
// textured = true adds the C4 rule: a quarter of the diffuse spheres alternate between a marble
// NoiseTexture(4) and a CheckerTexture(0.32) sharing one Perlin table.
void build_spheres(RefScene &s, int half, bool textured) {
  Tex ground_tex = s.checker(0.32, Color(.2, .3, .1), Color(.9, .9, .9));
  s.add_sphere(Point3(0, -1000, 0), 1000, s.lambertian(ground_tex));

  Mat marble{}, check{};
  int textured_count = 0;
  if (textured) {
    PerlinNoise perlin;
    int pidx = s.add_perlin(perlin);
    marble = s.lambertian(s.noise(4, perlin, pidx));
    check = s.lambertian(s.checker(0.32, Color(.8, .1, .1), Color(.9, .9, .9)));
  }

  for (int a = -half; a < half; a++) {
    for (int b = -half; b < half; b++) {
      double choose_mat = random_double();
      Point3 center(a + 0.9 * random_double(), 0.2, b + 0.9 * random_double());

      if ((center - Point3(4, 0.2, 0)).length() > 0.9) {
        if (choose_mat < 0.8) {
          Color albedo = Color::random() * Color::random();
          Point3 center2 = center + Vec3(0, random_double(0, .5), 0);
          Mat m;
          if (textured && random_double() < 0.25)
            m = (textured_count++ % 2 == 0) ? marble : check;
          else
            m = s.lambertian(albedo);
          s.add_moving_sphere(center, center2, 0.2, m);
        } else if (choose_mat < 0.95) {
          Color albedo = Color::random(0.5, 1);
          double fuzz = random_double(0, 0.5);
          s.add_sphere(center, 0.2, s.metal(albedo, fuzz));
        } else {
          s.add_sphere(center, 0.2, s.dielectric(1.5));
        }
      }
    }
  }

  s.add_sphere(Point3(0, 1, 0), 1.0, s.dielectric(1.5));
  s.add_sphere(Point3(-4, 1, 0), 1.0, s.lambertian(Color(0.4, 0.2, 0.1)));
  s.add_sphere(Point3(4, 1, 0), 1.0, s.metal(Color(0.7, 0.6, 0.5), 0.0));

  s.cam.aspect_ratio = 16.0 / 9.0;
  s.cam.background = Color(0.70, 0.80, 1.00);
  s.cam.vfov = 20;
  s.cam.lookfrom = Point3(13, 2, 3);
  s.cam.lookat = Point3(0, 0, 0);
  s.cam.vup = Vec3(0, 1, 0);
  s.cam.defocus_angle = 0.6;
  s.cam.focus_dist = 10.0;
}

void cornell_walls(RefScene &s) {
  Mat red = s.lambertian(Color(.65, .05, .05));
  Mat white = s.lambertian(Color(.73, .73, .73));
  Mat green = s.lambertian(Color(.12, .45, .15));
  Mat light = s.diffuse_light(Color(15, 15, 15));
  s.add_quad(Point3(555, 0, 0), Vec3(0, 0, 555), Vec3(0, 555, 0), green);
  s.add_quad(Point3(0, 0, 555), Vec3(0, 0, -555), Vec3(0, 555, 0), red);
  s.add_quad(Point3(0, 555, 0), Vec3(555, 0, 0), Vec3(0, 0, 555), white);
  s.add_quad(Point3(0, 0, 555), Vec3(555, 0, 0), Vec3(0, 0, -555), white);
  s.add_quad(Point3(555, 0, 555), Vec3(-555, 0, 0), Vec3(0, 555, 0), white);
  s.add_quad(Point3(213, 554, 227), Vec3(130, 0, 0), Vec3(0, 0, 105), light);

  s.cam.aspect_ratio = 1.0;
  s.cam.background = Color(0, 0, 0);
  s.cam.vfov = 40;
  s.cam.lookfrom = Point3(278, 278, -800);
  s.cam.lookat = Point3(278, 278, 0);
  s.cam.vup = Vec3(0, 1, 0);
  s.cam.defocus_angle = 0;
}

// populate_cornell_box_scene (main.cpp:21-71).
void build_cornell(RefScene &s) {
  cornell_walls(s);
  Mat white = s.lambertian(Color(.73, .73, .73));
  s.add_box(Point3(0, 0, 0), Point3(165, 330, 165), white, {translate(Vec3(265, 0, 295)), rotate_y(15)});
  s.add_sphere(Point3(190, 90, 190), 90, s.dielectric(1.5));
  s.add_light_quad(Point3(343, 554, 332), Vec3(-130, 0, 0), Vec3(0, 0, -105));
  s.add_light_sphere(Point3(190, 90, 190), 90);
}

// C3: Cornell box whose two blocks are constant-density smoke.
void build_cornell_smoke(RefScene &s) {
  cornell_walls(s);
  s.add_box_medium(Point3(0, 0, 0), Point3(165, 330, 165), {translate(Vec3(265, 0, 295)), rotate_y(15)}, 0.01,
                   Color(0, 0, 0));
  s.add_box_medium(Point3(0, 0, 0), Point3(165, 165, 165), {translate(Vec3(130, 0, 65)), rotate_y(-18)}, 0.01,
                   Color(1, 1, 1));
  s.add_light_quad(Point3(343, 554, 332), Vec3(-130, 0, 0), Vec3(0, 0, -105));
}

// C5: "final scene" assembled from reference classes.  boxes_per_side = 20 and
// n_cluster = 1000 give the full scene; smaller values give quick test variants.
void build_final(RefScene &s, int boxes_per_side, int n_cluster) {
  Mat ground = s.lambertian(Color(0.48, 0.83, 0.53));
  for (int i = 0; i < boxes_per_side; i++) {
    for (int j = 0; j < boxes_per_side; j++) {
      double w = 100.0;
      double x0 = -1000.0 + i * w;
      double z0 = -1000.0 + j * w;
      double y0 = 0.0;
      double x1 = x0 + w;
      double y1 = random_double(1, 101);
      double z1 = z0 + w;
      s.add_box(Point3(x0, y0, z0), Point3(x1, y1, z1), ground, {});
    }
  }
  Mat light = s.diffuse_light(Color(7, 7, 7));
  s.add_quad(Point3(123, 554, 147), Vec3(300, 0, 0), Vec3(0, 0, 265), light);
  s.add_light_quad(Point3(123, 554, 147), Vec3(300, 0, 0), Vec3(0, 0, 265));

  Point3 center1(400, 400, 200);
  Point3 center2 = center1 + Vec3(30, 0, 0);
  s.add_moving_sphere(center1, center2, 50, s.lambertian(Color(0.7, 0.3, 0.1)));
  s.add_sphere(Point3(260, 150, 45), 50, s.dielectric(1.5));
  s.add_sphere(Point3(0, 150, 145), 50, s.metal(Color(0.8, 0.8, 0.9), 1.0));

  s.add_sphere(Point3(360, 150, 145), 70, s.dielectric(1.5));
  s.add_sphere_medium(Point3(360, 150, 145), 70, 0.2, Color(0.2, 0.4, 0.9));
  s.add_sphere_medium(Point3(0, 0, 0), 5000, 0.0001, Color(1, 1, 1));

  // The reference has no image texture: a checker sphere stands in for the globe.
  s.add_sphere(Point3(400, 200, 400), 100, s.lambertian(s.checker(20.0, Color(.1, .2, .7), Color(.9, .9, .9))));
  PerlinNoise perlin;
  int pidx = s.add_perlin(perlin);
  s.add_sphere(Point3(220, 280, 300), 80, s.lambertian(s.noise(0.2, perlin, pidx)));

  Mat white = s.lambertian(Color(.73, .73, .73));
  int xf = s.add_xform({translate(Vec3(-100, 270, 395)), rotate_y(15)});
  std::vector<XF> chain = {translate(Vec3(-100, 270, 395)), rotate_y(15)};
  for (int j = 0; j < n_cluster; j++) {
    Point3 c = Point3::random(0, 165);
    s.add_object(RefScene::wrap(s.make_sphere(c, c, false, 10, white, xf, s.n_objects, 0), chain));
  }

  s.cam.aspect_ratio = 16.0 / 9.0;
  s.cam.background = Color(0, 0, 0);
  s.cam.vfov = 40;
  s.cam.lookfrom = Point3(478, 278, -600);
  s.cam.lookat = Point3(278, 278, 0);
  s.cam.vup = Vec3(0, 1, 0);
  s.cam.defocus_angle = 0;
}

// Median-split tree out of reference BVHNode(left,right) nodes (never flattened).
HittablePtr median_tree(std::vector<HittablePtr> &objs, size_t lo, size_t hi) {
  if (hi - lo == 1)
    return objs[lo];
  AABB box = objs[lo]->get_bounding_box();
  for (size_t i = lo + 1; i < hi; i++)
    box = AABB(box, objs[i]->get_bounding_box());
  int axis = box.get_longest_axis();
  size_t mid = lo + (hi - lo) / 2;
  std::nth_element(objs.begin() + lo, objs.begin() + mid, objs.begin() + hi,
                   [axis](const HittablePtr &a, const HittablePtr &b) {
                     return a->get_bounding_box().center()[axis] < b->get_bounding_box().center()[axis];
                   });
  HittablePtr l = median_tree(objs, lo, mid);
  HittablePtr r = median_tree(objs, mid, hi);
  return std::make_shared<BVHNode>(l, r);
}

std::shared_ptr<HittableList> accelerate(const HittableList &list) {
  if (list.get_objects().empty())
    return std::make_shared<HittableList>();
  if (list.get_objects().size() <= 400) // reference build + flattened traversal (StaticCamera.cpp:35-40)
    return std::make_shared<HittableList>(std::make_shared<BVHNode>(list));
  std::vector<HittablePtr> objs = list.get_objects();
  return std::make_shared<HittableList>(median_tree(objs, 0, objs.size()));
}

class HarnessCamera : public Camera {
public:
  explicit HarnessCamera(const CameraConfig &c) : Camera(c) { initialize(); }
  void render(HittableList &, HittableList &) override {}
  Ray ray(int i, int j, int s_i, int s_j) const { return get_ray(i, j, s_i, s_j); }
  Color color(const Ray &r, int depth, const HittableList &w, const HittableList &l) const {
    return ray_color(r, depth, w, l);
  }
  int height() const { return m_image_height; }
  int width() const { return m_image_width; }
  double scale() const { return m_pixel_samples_scale; }
  void export_derived(rt_camera *out) const {
    out->image_width = m_image_width;
    out->image_height = m_image_height;
    set3(out->center, m_center);
    set3(out->pixel00_loc, m_pixel00_loc);
    set3(out->pixel_delta_u, m_pixel_delta_u);
    set3(out->pixel_delta_v, m_pixel_delta_v);
    set3(out->defocus_disk_u, m_defocus_disk_u);
    set3(out->defocus_disk_v, m_defocus_disk_v);
    out->defocus_angle = m_defocus_angle;
    set3(out->background, m_background);
  }
};

CameraConfig make_config(const RefScene &s, int width, int spp, int depth) {
  CameraConfig c = s.cam;
  c.image_width = width;
  c.samples_per_pixel = spp;
  c.max_depth = depth;
  return c;
}

const HittableList &pick_world(RefScene *s, int use_bvh) {
  if (!use_bvh)
    return s->world;
  if (!s->bvh_world)
    s->bvh_world = accelerate(s->world);
  return *s->bvh_world;
}
const HittableList &pick_lights(RefScene *s, int /*use_bvh*/) {
  // StaticCamera.cpp:38-39 also wraps the lights in a BVHNode.  For one or two lights (all fixtures)
  // BVHNode::pdf_value/random (BVHNode.cpp:149-166) give the same values and consume the same number
  // of engine calls as HittableList's (HittableList.cpp:44-63); for more lights the BVH version
  // selects lights non-uniformly while averaging their pdfs uniformly, which is not reproduced.
  return s->lights;
}

} // namespace

extern "C" {

// name: "spheres" (p0 = grid half size, default 11), "spheres_textured" (C4 generator, p0 = half
// size), "cornell", "cornell_smoke", "final" (p0 = boxes per side, p1 = cluster spheres).
static void *scene_build(const char *name, uint64_t seed, int p0, int p1, bool plain);
void *ref_scene_build(const char *name, uint64_t seed, int p0, int p1) { return scene_build(name, seed, p0, p1, false); }

static void *scene_build(const char *name, uint64_t seed, int p0, int p1, bool plain) {
  random_engine().seed(static_cast<std::mt19937::result_type>(seed));
  RefScene *s = new RefScene();
  s->plain = plain;
  std::string n(name);
  if (n == "spheres")
    build_spheres(*s, p0 > 0 ? p0 : 11, false);
  else if (n == "spheres_textured")
    build_spheres(*s, p0 > 0 ? p0 : 11, true);
  else if (n == "cornell")
    build_cornell(*s);
  else if (n == "cornell_smoke")
    build_cornell_smoke(*s);
  else if (n == "final")
    build_final(*s, p0 > 0 ? p0 : 20, p1 >= 0 ? p1 : 1000);
  else {
    delete s;
    return nullptr;
  }
  s->finalize();
  return s;
}

void ref_scene_free(void *h) { delete static_cast<RefScene *>(h); }

const rt_scene_desc *ref_scene_desc(void *h) { return &static_cast<RefScene *>(h)->desc; }

// The scene's camera (CameraConfig fields the scene function sets, main.cpp:64-70,123-130) with the
// CLI-controlled width / spp / depth filled in by the caller.
void ref_scene_camera_config(void *h, int width, int spp, int depth, rt_camera_config *out) {
  RefScene *s = static_cast<RefScene *>(h);
  CameraConfig c = make_config(*s, width, spp, depth);
  std::memset(out, 0, sizeof *out);
  out->image_width = c.image_width;
  out->samples_per_pixel = c.samples_per_pixel;
  out->max_depth = c.max_depth;
  out->aspect_ratio = c.aspect_ratio;
  out->vfov = c.vfov;
  out->defocus_angle = c.defocus_angle;
  out->focus_dist = c.focus_dist;
  set3(out->lookfrom, c.lookfrom);
  set3(out->lookat, c.lookat);
  set3(out->vup, c.vup);
  set3(out->background, c.background);
}

// The reference's CLI forces one aspect ratio on every scene (BASELINE config 3 renders the Cornell box at 16:9,
// its scene function sets 1.0): override the scene camera's aspect ratio for the renders that follow.
void ref_scene_set_aspect(void *h, double aspect_ratio) { static_cast<RefScene *>(h)->cam.aspect_ratio = aspect_ratio; }

// Camera::initialize through the reference (for pinning rt_camera_init).
void ref_camera_init(const rt_camera_config *cfg, rt_camera *out) {
  CameraConfig c;
  c.image_width = cfg->image_width;
  c.samples_per_pixel = cfg->samples_per_pixel;
  c.max_depth = cfg->max_depth;
  c.aspect_ratio = cfg->aspect_ratio;
  c.vfov = cfg->vfov;
  c.defocus_angle = cfg->defocus_angle;
  c.focus_dist = cfg->focus_dist;
  c.lookfrom = Point3(cfg->lookfrom[0], cfg->lookfrom[1], cfg->lookfrom[2]);
  c.lookat = Point3(cfg->lookat[0], cfg->lookat[1], cfg->lookat[2]);
  c.vup = Vec3(cfg->vup[0], cfg->vup[1], cfg->vup[2]);
  c.background = Color(cfg->background[0], cfg->background[1], cfg->background[2]);
  HarnessCamera cam(c);
  cam.export_derived(out);
}

// Camera::get_ray for every pixel (row-major), stratum (s_i, s_j), after seeding the engine.
// Fills origin/direction/time; t_min/t_max are set to the reference's (0.001, inf).
void ref_primary_rays(void *h, int width, int spp, uint64_t seed, int s_i, int s_j, rt_ray *out) {
  RefScene *s = static_cast<RefScene *>(h);
  HarnessCamera cam(make_config(*s, width, spp, 1));
  random_engine().seed(static_cast<std::mt19937::result_type>(seed));
  int W = cam.width(), H = cam.height();
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      Ray r = cam.ray(i, j, s_i, s_j);
      rt_ray &o = out[size_t(j) * W + i];
      std::memset(&o, 0, sizeof o);
      set3(o.origin, r.origin());
      set3(o.direction, r.direction());
      o.time = r.time();
      o.t_min = 0.001;
      o.t_max = INF;
      o.rng_pixel = uint32_t(j) * uint32_t(W) + uint32_t(i);
    }
}

// Closest hit through the reference's Hittable::hit.  use_bvh = 0: HittableList linear scan
// (HittableList.cpp:26-42); 1: the accelerated world.  Scenes with media consume the thread's RNG.
void ref_trace(void *h, const rt_ray *rays, int64_t n, int use_bvh, rt_hit *hits) {
  RefScene *s = static_cast<RefScene *>(h);
  const HittableList &world = pick_world(s, use_bvh);
  for (int64_t k = 0; k < n; k++) {
    const rt_ray &q = rays[k];
    Ray r(Point3(q.origin[0], q.origin[1], q.origin[2]), Vec3(q.direction[0], q.direction[1], q.direction[2]),
          q.time);
    HitRecord rec;
    tl_last_object = -1;
    bool ok = world.hit(r, Interval(q.t_min, q.t_max), rec);
    hits[k].t = ok ? rec.t : INF;
    hits[k].prim = -1; // the reference has no primitive ids
    hits[k].object = ok ? tl_last_object : -1;
    hits[k].front_face = ok ? int(rec.frontFace) : 0;
    hits[k].pad_ = 0;
  }
}

void ref_seed(uint64_t seed) { random_engine().seed(static_cast<std::mt19937::result_type>(seed)); }
double ref_random_double(void) { return random_double(); }
int ref_random_int(int lo, int hi) { return random_int(lo, hi); }

// Renders rows [row0,row1) of the image exactly as StaticCamera::render_cpu's serial loop does
// (StaticCamera.cpp:101-131): sum over sqrt_spp^2 strata, times 1/spp.  out = (row1-row0)*W*3 doubles.
//   n_threads <= 1: on the calling thread after random_engine().seed(seed) -> deterministic.
//   n_threads  > 1: scanlines handed to n_threads std::threads (thread t seeds seed+1+t).
//   single_stratum >= 0: dynamic-mode frame (DynamicCamera.cpp:103-171): only stratum
//     (s % sqrt_spp, s / sqrt_spp), un-normalised.
// Returns wall seconds of the render loop; *segments = number of world.hit calls.
double ref_render(void *h, int width, int spp, int depth, uint64_t seed, int use_bvh, int n_threads, int row0,
                  int row1, int single_stratum, double *out, uint64_t *segments) {
  RefScene *s = static_cast<RefScene *>(h);
  HarnessCamera cam(make_config(*s, width, spp, depth));
  const HittableList &accel = pick_world(s, use_bvh);
  const HittableList &lights = pick_lights(s, use_bvh);
  HittableList world(std::make_shared<SegmentCounter>(std::make_shared<HittableList>(accel)));
  int W = cam.width(), H = cam.height();
  if (row1 > H || row1 < 0)
    row1 = H;
  int sqrt_spp = int(std::sqrt(double(spp)));
  double scale = single_stratum >= 0 ? 1.0 : cam.scale();

  auto render_row = [&](int j) {
    for (int i = 0; i < W; i++) {
      Color pixel(0, 0, 0);
      if (single_stratum >= 0) {
        Ray r = cam.ray(i, j, single_stratum % sqrt_spp, single_stratum / sqrt_spp);
        pixel += cam.color(r, depth, world, lights);
      } else {
        for (int s_j = 0; s_j < sqrt_spp; ++s_j)
          for (int s_i = 0; s_i < sqrt_spp; ++s_i) {
            Ray r = cam.ray(i, j, s_i, s_j);
            pixel += cam.color(r, depth, world, lights);
          }
      }
      Color c = scale * pixel;
      double *o = out + (size_t(j - row0) * W + i) * 3;
      o[0] = c.x();
      o[1] = c.y();
      o[2] = c.z();
    }
  };

  std::atomic<uint64_t> total_segments{0};
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) {
    random_engine().seed(static_cast<std::mt19937::result_type>(seed));
    tl_segments = 0;
    for (int j = row0; j < row1; j++)
      render_row(j);
    total_segments += tl_segments;
  } else {
    std::atomic<int> next{row0};
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++)
      pool.emplace_back([&, t]() {
        random_engine().seed(static_cast<std::mt19937::result_type>(seed + 1 + t));
        tl_segments = 0;
        for (;;) {
          int j = next.fetch_add(1);
          if (j >= row1)
            break;
          render_row(j);
        }
        total_segments += tl_segments;
      });
    for (std::thread &t : pool)
      t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (segments)
    *segments = total_segments.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

#ifdef USE_CUDA
// ---------------------------------------------------------------------------------------------------
// The reference's own GPU path, timed as an informational comparator (oracle/_ref_gpu/libref_gpu.so: the
// reference's .cu and .cpp sources compiled with -DUSE_CUDA for sm_100, nothing of this repository's library).
// One dynamic-mode frame exactly as DynamicCamera::render_gpu drives it (DynamicCamera.cpp:434-554): scene
// converted by initialize_cuda_scene, one curand state per pixel, then per displayed frame one
// cuda_dynamic_render_tile_wrapper launch + cudaDeviceSynchronize per tile and a full-buffer copy to the host.
//   tile_size > 0 : the reference's loop (its default tile size is 32)
//   tile_size = 0 : ONE launch of the same kernel over the whole frame (what the kernel itself can do)
// Returns 0 on success.  ms_frame = wall time per frame of the launch loop (+ the device-to-host copy when
// with_copy), mean = mean of the accumulated radiance / frames (a sanity check of the picture).
// ---------------------------------------------------------------------------------------------------
int ref_gpu_frame(const char *name, uint64_t seed, int p0, int p1, int width, int depth, int tile_size, int frames,
                  int with_copy, double *ms_frame, double *mean, char *error, int error_len) {
  auto fail = [&](const std::string &msg) {
    if (error && error_len > 0)
      std::snprintf(error, error_len, "%s", msg.c_str());
    return 1;
  };
  // The reference never raises the device stack limit; its recursive ray_color_cuda (CameraKernels.cu:106-202)
  // overruns the 1 KB default at depth 8 on this toolchain ("illegal memory access").  The harness - not the
  // reference sources - raises it so that the kernels can be timed at all.
  {
    const char *e = std::getenv("REF_GPU_STACK");
    cudaDeviceSetLimit(cudaLimitStackSize, e ? (size_t)std::atol(e) : 32768);
  }
  RefScene *s = static_cast<RefScene *>(scene_build(name, seed, p0, p1, true));
  if (!s)
    return fail("unknown scene");
  HarnessCamera cam(make_config(*s, width, 1, depth));
  rt_camera rc{};
  cam.export_derived(&rc);
  const int W = cam.width(), H = cam.height();
  // -b: the world (and the lights) wrapped in the reference's BVH (StaticCamera.cpp:140-145)
  HittableList world = s->world, lights = s->lights;
  if (!std::getenv("REF_GPU_NO_BVH")) { // REF_GPU_NO_BVH: the list world (no -b), to tell a traversal fault from the rest
    if (!world.get_objects().empty())
      world = HittableList(std::make_shared<BVHNode>(world));
    if (!lights.get_objects().empty())
      lights = HittableList(std::make_shared<BVHNode>(lights));
  }

  CudaColor *d_accum = nullptr;
  curandState *d_rand = nullptr;
  const size_t accum_bytes = size_t(W) * H * sizeof(CudaColor);
  if (cudaMalloc(&d_accum, accum_bytes) != cudaSuccess || cudaMalloc(&d_rand, size_t(W) * H * sizeof(curandState)) != cudaSuccess)
    return fail("cudaMalloc failed");
  cudaMemset(d_accum, 0, accum_bytes);
  dim3 block(16, 16), grid((W + 15) / 16, (H + 15) / 16);
  cuda_init_rand_states_wrapper(d_rand, W, H, (unsigned long)seed, grid, block);
  if (cudaDeviceSynchronize() != cudaSuccess)
    return fail(std::string("init_rand_states: ") + cudaGetErrorString(cudaGetLastError()));
  CudaSceneData scene = initialize_cuda_scene(world, lights, false);
  if (!scene.world.get() || !scene.lights.get())
    return fail("initialize_cuda_scene returned null");

  auto v3 = [](const double *p) { return CudaVec3(p[0], p[1], p[2]); };
  // u, v, w only feed the kernel's signature (get_ray_cuda uses the derived vectors); recomputed as Camera::initialize does
  Vec3 cw = unit_vector(s->cam.lookfrom - s->cam.lookat), cu = unit_vector(cross_product(s->cam.vup, cw)), cv = cross_product(cw, cu);
  auto frame = [&](int f) {
    const int ts = tile_size > 0 ? tile_size : std::max(W, H);
    const int ntx = (W + ts - 1) / ts, nty = (H + ts - 1) / ts;
    for (int tile = 0; tile < ntx * nty; ++tile) {
      int sr = (tile / ntx) * ts, sc = (tile % ntx) * ts;
      int er = std::min(sr + ts, H), ec = std::min(sc + ts, W);
      dim3 tb(16, 16), tg((ec - sc + 15) / 16, (er - sr + 15) / 16);
      cuda_dynamic_render_tile_wrapper(d_accum, W, H, sr, er, sc, ec, 0, 0, 1, depth, v3(rc.center), v3(rc.pixel00_loc),
                                       v3(rc.pixel_delta_u), v3(rc.pixel_delta_v), CudaVec3(cu.x(), cu.y(), cu.z()),
                                       CudaVec3(cv.x(), cv.y(), cv.z()), CudaVec3(cw.x(), cw.y(), cw.z()),
                                       v3(rc.defocus_disk_u), v3(rc.defocus_disk_v), rc.defocus_angle, v3(rc.background),
                                       scene.world.get(), scene.lights.get(), d_rand, tg, tb);
      cudaDeviceSynchronize();
    }
    (void)f;
  };
  std::vector<CudaColor> host(size_t(W) * H);
  frame(-1); // warm-up
  if (cudaDeviceSynchronize() != cudaSuccess) {
    std::string msg = std::string("render kernel: ") + cudaGetErrorString(cudaGetLastError());
    scene.world.release(); // the context is gone: the reference's cudaFree wrappers would exit(1) on it
    scene.lights.release();
    return fail(msg);
  }
  cudaMemset(d_accum, 0, accum_bytes);
  auto t0 = std::chrono::steady_clock::now();
  for (int f = 0; f < frames; f++) {
    frame(f);
    if (with_copy)
      cudaMemcpy(host.data(), d_accum, accum_bytes, cudaMemcpyDeviceToHost);
  }
  cudaError_t e = cudaDeviceSynchronize();
  auto t1 = std::chrono::steady_clock::now();
  if (e != cudaSuccess)
    return fail(std::string("render loop: ") + cudaGetErrorString(e));
  cudaMemcpy(host.data(), d_accum, accum_bytes, cudaMemcpyDeviceToHost);
  double sum = 0;
  for (const CudaColor &c : host) {
    double v = (c.x + c.y + c.z) / 3.0;
    sum += std::isfinite(v) ? v : 0.0;
  }
  if (ms_frame)
    *ms_frame = std::chrono::duration<double, std::milli>(t1 - t0).count() / std::max(frames, 1);
  if (mean)
    *mean = sum / (double(W) * H) / std::max(frames, 1);
  cleanup_cuda_scene(scene);
  cudaFree(d_accum);
  cudaFree(d_rand);
  delete s;
  return 0;
}
#endif // USE_CUDA

// to_byte (ColorUtility.hpp:18-23) for pinning the tonemap.
int ref_to_byte(double v) { return int(to_byte(v)); }

int ref_hardware_threads(void) { return int(std::thread::hardware_concurrency()); }

} // extern "C"
