// Host-side scene construction: fills the flat arrays of rt_scene_desc (include/rt_b200.h).
// The vocabulary follows the reference's scene classes (spheres, quads, boxes, translate / rotate_y
// instances, constant media; lambertian / metal / dielectric / diffuse_light / isotropic materials;
// solid / checker / noise textures) but nothing here is a class hierarchy: every call appends a
// record to a flat array.
#pragma once

#include "../../include/rt_b200.h"

#include <cmath>
#include <random>
#include <string>
#include <vector>

namespace rth {

struct Vec {
  double x = 0, y = 0, z = 0;
  Vec() = default;
  Vec(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
};
inline Vec operator+(Vec a, Vec b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec operator-(Vec a, Vec b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec operator-(Vec a) { return {-a.x, -a.y, -a.z}; }
inline void store(double *dst, Vec v) {
  dst[0] = v.x;
  dst[1] = v.y;
  dst[2] = v.z;
}

// The reference draws from one std::mt19937 through std::uniform_real_distribution /
// std::uniform_int_distribution (utils/math/Utility.hpp:16-37); using the same standard-library
// types gives the same scenes for the same seed.
struct Rng {
  std::mt19937 engine;
  explicit Rng(uint64_t seed) : engine(static_cast<std::mt19937::result_type>(seed)) {}
  double uniform() {
    std::uniform_real_distribution<double> d(0.0, 1.0);
    return d(engine);
  }
  double uniform(double lo, double hi) {
    std::uniform_real_distribution<double> d(lo, hi);
    return d(engine);
  }
  int integer(int lo, int hi) {
    std::uniform_int_distribution<int> d(lo, hi);
    return d(engine);
  }
  // Vec3::random (utils/math/Vec3.hpp:108-115): as compiled, the z draw comes first.
  Vec vec() {
    double z = uniform(), y = uniform(), x = uniform();
    return {x, y, z};
  }
  Vec vec(double lo, double hi) {
    double z = uniform(lo, hi), y = uniform(lo, hi), x = uniform(lo, hi);
    return {x, y, z};
  }
};

struct XformOp {
  int type;
  Vec offset;
  double angle_deg;
};
inline XformOp translate(Vec o) { return {RT_XF_TRANSLATE, o, 0.0}; }
inline XformOp rotate_y(double deg) { return {RT_XF_ROTATE_Y, Vec(), deg}; }

class SceneBuilder {
public:
  rt_camera_config camera{}; // scene-specific fields; width / spp / depth come from the CLI

  // textures -> index
  int solid(Vec color);
  int checker(double scale, Vec even, Vec odd);
  int perlin(Rng &rng); // a fresh Perlin table (utils/math/PerlinNoise.hpp:19-26)
  int noise(double scale, int perlin_index);
  // image textures (not in the reference; include/rt_b200.h rt_image): rgb = width * height * 3 bytes
  int image(int width, int height, const std::vector<uint8_t> &rgb);
  int image_texture(int image_index);
  // materials -> index
  int lambertian(Vec albedo) { return lambertian_tex(solid(albedo)); }
  int lambertian_tex(int texture);
  int metal(Vec albedo, double fuzz);
  int dielectric(double ior);
  int diffuse_light(Vec emit);
  int isotropic(Vec albedo);
  // instancing chains (outermost wrapper first) -> index, -1 for an empty chain
  int xform(const std::vector<XformOp> &chain);
  // top-level world objects
  void sphere(Vec center, double radius, int material, int xf = -1);
  void moving_sphere(Vec center0, Vec center1, double radius, int material);
  void quad(Vec corner, Vec u, Vec v, int material);
  void box(Vec a, Vec b, int material, const std::vector<XformOp> &chain);
  void box_medium(Vec a, Vec b, const std::vector<XformOp> &chain, double density, Vec albedo);
  void sphere_medium(Vec center, double radius, double density, Vec albedo);
  // light-sampling proxies
  void light_quad(Vec corner, Vec u, Vec v);
  void light_sphere(Vec center, double radius);

  // raw access (JSON loader)
  std::vector<rt_sphere> spheres;
  std::vector<rt_quad> quads;
  std::vector<rt_xform_op> xform_ops;
  std::vector<rt_xform> xforms;
  std::vector<rt_medium> media;
  std::vector<rt_material> materials;
  std::vector<rt_texture> textures;
  std::vector<rt_perlin> perlins;
  std::vector<rt_light> lights;
  std::vector<rt_image> images;                 // rgb pointers are refreshed by finalize()
  std::vector<std::vector<uint8_t>> image_data; // owned texel bytes, one vector per image
  int n_objects = 0;

  const rt_scene_desc *finalize();

private:
  void box_sides(Vec a, Vec b, int material, int xf, int object, int flags);
  rt_scene_desc m_desc{};
};

// Built-in scenes (scenes.cpp).
bool build_builtin(SceneBuilder &s, const std::string &name, uint64_t seed, int p0, int p1);

} // namespace rth

struct rth_scene {
  rth::SceneBuilder builder;
};
