#include "scene_builder.hpp"

#include <algorithm>
#include <cstring>

namespace rth {

namespace {
const double kPi = 3.1415926535897932385; // utils/math/Utility.hpp:8
}

int SceneBuilder::solid(Vec color) {
  rt_texture t{};
  t.type = RT_TEX_SOLID;
  t.even = t.odd = t.perlin = -1;
  store(t.color, color);
  textures.push_back(t);
  return int(textures.size()) - 1;
}

int SceneBuilder::checker(double scale, Vec even, Vec odd) {
  int e = solid(even), o = solid(odd);
  rt_texture t{};
  t.type = RT_TEX_CHECKER;
  t.even = e;
  t.odd = o;
  t.perlin = -1;
  t.scale = scale;
  textures.push_back(t);
  return int(textures.size()) - 1;
}

// PerlinNoise::PerlinNoise (utils/math/PerlinNoise.hpp:19-26,162-178): 256 unit gradients, then three
// Fisher-Yates permutations, all from the scene's generator.
int SceneBuilder::perlin(Rng &rng) {
  rt_perlin p{};
  for (int i = 0; i < RT_PERLIN_POINTS; i++) {
    Vec v = rng.vec(-1, 1);
    double len = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    if (len > 1e-8) { // Vec3::normalize (utils/math/Vec3.hpp:141-149)
      double s = 1.0 / len;
      v = Vec(v.x * s, v.y * s, v.z * s);
    } else {
      v = Vec(1.0, 0.0, 0.0);
    }
    store(p.rand_vec[i], v);
  }
  for (int32_t *perm : {p.perm_x, p.perm_y, p.perm_z}) {
    for (int i = 0; i < RT_PERLIN_POINTS; i++)
      perm[i] = i;
    for (int i = RT_PERLIN_POINTS - 1; i > 0; i--)
      std::swap(perm[i], perm[rng.integer(0, i)]);
  }
  perlins.push_back(p);
  return int(perlins.size()) - 1;
}

int SceneBuilder::noise(double scale, int perlin_index) {
  rt_texture t{};
  t.type = RT_TEX_NOISE;
  t.even = t.odd = -1;
  t.perlin = perlin_index;
  t.scale = scale;
  textures.push_back(t);
  return int(textures.size()) - 1;
}

int SceneBuilder::image(int width, int height, const std::vector<uint8_t> &rgb) {
  rt_image im{};
  im.width = width;
  im.height = height;
  images.push_back(im);
  image_data.push_back(rgb);
  image_data.back().resize(size_t(width > 0 ? width : 0) * size_t(height > 0 ? height : 0) * 3);
  return int(images.size()) - 1;
}

int SceneBuilder::image_texture(int image_index) {
  rt_texture t{};
  t.type = RT_TEX_IMAGE;
  t.even = t.odd = -1;
  t.perlin = image_index; // rt_texture: the table index field doubles as the image index
  textures.push_back(t);
  return int(textures.size()) - 1;
}

static int push_material(std::vector<rt_material> &v, int type, int tex, Vec albedo, double fuzz, double ior) {
  rt_material m{};
  m.type = type;
  m.texture = tex;
  store(m.albedo, albedo);
  m.fuzz = fuzz;
  m.ior = ior;
  v.push_back(m);
  return int(v.size()) - 1;
}

int SceneBuilder::lambertian_tex(int texture) {
  return push_material(materials, RT_MAT_LAMBERTIAN, texture, Vec(), 0, 0);
}
int SceneBuilder::metal(Vec albedo, double fuzz) {
  return push_material(materials, RT_MAT_METAL, -1, albedo, fuzz, 0);
}
int SceneBuilder::dielectric(double ior) { return push_material(materials, RT_MAT_DIELECTRIC, -1, Vec(), 0, ior); }
int SceneBuilder::diffuse_light(Vec emit) {
  return push_material(materials, RT_MAT_DIFFUSE_LIGHT, solid(emit), Vec(), 0, 0);
}
int SceneBuilder::isotropic(Vec albedo) {
  return push_material(materials, RT_MAT_ISOTROPIC, solid(albedo), Vec(), 0, 0);
}

int SceneBuilder::xform(const std::vector<XformOp> &chain) {
  if (chain.empty())
    return -1;
  rt_xform x{int(xform_ops.size()), int(chain.size())};
  for (const XformOp &c : chain) {
    rt_xform_op op{};
    op.type = c.type;
    if (c.type == RT_XF_TRANSLATE) {
      store(op.offset, c.offset);
    } else { // RotateY::RotateY (scene/objects/RotateY.cpp:5-8)
      op.angle_deg = c.angle_deg;
      double radians = c.angle_deg * kPi / 180.0;
      op.sin_theta = std::sin(radians);
      op.cos_theta = std::cos(radians);
    }
    xform_ops.push_back(op);
  }
  xforms.push_back(x);
  return int(xforms.size()) - 1;
}

static rt_sphere make_sphere(Vec c0, Vec dir, double r, int material, int xf, int object, int flags) {
  rt_sphere s{};
  store(s.center0, c0);
  store(s.center_dir, dir);
  s.radius = std::fmax(0, r);
  s.material = material;
  s.xform = xf;
  s.object = object;
  s.flags = flags;
  return s;
}

void SceneBuilder::sphere(Vec center, double radius, int material, int xf) {
  spheres.push_back(make_sphere(center, Vec(), radius, material, xf, n_objects, 0));
  ++n_objects;
}

void SceneBuilder::moving_sphere(Vec center0, Vec center1, double radius, int material) {
  spheres.push_back(make_sphere(center0, center1 - center0, radius, material, -1, n_objects, 0));
  ++n_objects;
}

static rt_quad make_quad(Vec corner, Vec u, Vec v, int material, int xf, int object, int flags) {
  rt_quad q{};
  store(q.corner, corner);
  store(q.u, u);
  store(q.v, v);
  q.material = material;
  q.xform = xf;
  q.object = object;
  q.flags = flags;
  return q;
}

void SceneBuilder::quad(Vec corner, Vec u, Vec v, int material) {
  quads.push_back(make_quad(corner, u, v, material, -1, n_objects, 0));
  ++n_objects;
}

// The six sides of an axis-aligned box in the order make_box lists them
// (scene/objects/PlaneUtility.hpp:11-40): front, right, back, left, top, bottom.
void SceneBuilder::box_sides(Vec a, Vec b, int material, int xf, int object, int flags) {
  Vec lo(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z));
  Vec hi(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z));
  Vec dx(hi.x - lo.x, 0, 0), dy(0, hi.y - lo.y, 0), dz(0, 0, hi.z - lo.z);
  quads.push_back(make_quad(Vec(lo.x, lo.y, hi.z), dx, dy, material, xf, object, flags));
  quads.push_back(make_quad(Vec(hi.x, lo.y, hi.z), -dz, dy, material, xf, object, flags));
  quads.push_back(make_quad(Vec(hi.x, lo.y, lo.z), -dx, dy, material, xf, object, flags));
  quads.push_back(make_quad(Vec(lo.x, lo.y, lo.z), dz, dy, material, xf, object, flags));
  quads.push_back(make_quad(Vec(lo.x, hi.y, hi.z), dx, -dz, material, xf, object, flags));
  quads.push_back(make_quad(Vec(lo.x, lo.y, lo.z), dx, dz, material, xf, object, flags));
}

void SceneBuilder::box(Vec a, Vec b, int material, const std::vector<XformOp> &chain) {
  int xf = xform(chain);
  box_sides(a, b, material, xf, n_objects, 0);
  ++n_objects;
}

void SceneBuilder::box_medium(Vec a, Vec b, const std::vector<XformOp> &chain, double density, Vec albedo) {
  int xf = xform(chain);
  int first = int(quads.size());
  box_sides(a, b, -1, xf, n_objects, RT_PRIM_BOUNDARY);
  rt_medium m{};
  m.density = density;
  m.shape = RT_SHAPE_QUAD;
  m.first_prim = first;
  m.n_prims = 6;
  m.material = isotropic(albedo);
  m.object = n_objects;
  media.push_back(m);
  ++n_objects;
}

void SceneBuilder::sphere_medium(Vec center, double radius, double density, Vec albedo) {
  int first = int(spheres.size());
  spheres.push_back(make_sphere(center, Vec(), radius, -1, -1, n_objects, RT_PRIM_BOUNDARY));
  rt_medium m{};
  m.density = density;
  m.shape = RT_SHAPE_SPHERE;
  m.first_prim = first;
  m.n_prims = 1;
  m.material = isotropic(albedo);
  m.object = n_objects;
  media.push_back(m);
  ++n_objects;
}

void SceneBuilder::light_quad(Vec corner, Vec u, Vec v) {
  rt_light l{};
  l.shape = RT_SHAPE_QUAD;
  l.xform = -1;
  store(l.a, corner);
  store(l.b, u);
  store(l.c, v);
  lights.push_back(l);
}

void SceneBuilder::light_sphere(Vec center, double radius) {
  rt_light l{};
  l.shape = RT_SHAPE_SPHERE;
  l.xform = -1;
  store(l.a, center);
  l.radius = radius;
  lights.push_back(l);
}

const rt_scene_desc *SceneBuilder::finalize() {
  std::memset(&m_desc, 0, sizeof m_desc);
  m_desc.n_spheres = int(spheres.size());
  m_desc.n_quads = int(quads.size());
  m_desc.n_xform_ops = int(xform_ops.size());
  m_desc.n_xforms = int(xforms.size());
  m_desc.n_media = int(media.size());
  m_desc.n_materials = int(materials.size());
  m_desc.n_textures = int(textures.size());
  m_desc.n_perlins = int(perlins.size());
  m_desc.n_lights = int(lights.size());
  m_desc.n_objects = n_objects;
  m_desc.spheres = spheres.data();
  m_desc.quads = quads.data();
  m_desc.xform_ops = xform_ops.data();
  m_desc.xforms = xforms.data();
  m_desc.media = media.data();
  m_desc.materials = materials.data();
  m_desc.textures = textures.data();
  m_desc.perlins = perlins.data();
  m_desc.lights = lights.data();
  for (size_t i = 0; i < images.size(); i++)
    images[i].rgb = image_data[i].data();
  m_desc.n_images = int(images.size());
  m_desc.images = images.data();
  return &m_desc;
}

} // namespace rth
