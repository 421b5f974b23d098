// Built-in scenes.  The reference has no scene files: its scenes are two C++ functions in main.cpp
// (main.cpp:21-131).  These builders produce the same objects, in the same order, from the same
// std::mt19937 stream, as flat records; the remaining scenes are the benchmark configurations of
// BASELINE.json (Cornell box with smoke, million-sphere field, "final" scene).
#include "scene_builder.hpp"

namespace rth {

namespace {

void set_camera(SceneBuilder &s, double aspect, Vec background, double vfov, Vec from, Vec at, double defocus_angle,
                double focus_dist) {
  s.camera.aspect_ratio = aspect;
  store(s.camera.background, background);
  s.camera.vfov = vfov;
  store(s.camera.lookfrom, from);
  store(s.camera.lookat, at);
  store(s.camera.vup, Vec(0, 1, 0));
  s.camera.defocus_angle = defocus_angle;
  s.camera.focus_dist = focus_dist;
}

// populate_bouncing_spheres_scene (main.cpp:73-131).  `half` = 11 is the reference's scene; the
// million-sphere field of config 4 is the same loop with half = 500.  With `textured`, a quarter of the
// diffuse spheres alternate between a marble noise texture and a checker sharing one Perlin table.
void spheres_scene(SceneBuilder &s, Rng &rng, int half, bool textured) {
  s.sphere(Vec(0, -1000, 0), 1000, s.lambertian_tex(s.checker(0.32, Vec(.2, .3, .1), Vec(.9, .9, .9))));

  int marble = -1, check = -1, textured_count = 0;
  if (textured) {
    int table = s.perlin(rng);
    marble = s.lambertian_tex(s.noise(4, table));
    check = s.lambertian_tex(s.checker(0.32, Vec(.8, .1, .1), Vec(.9, .9, .9)));
  }

  for (int a = -half; a < half; a++) {
    for (int b = -half; b < half; b++) {
      double choose_mat = rng.uniform();
      // the compiled reference draws the z offset before the x offset
      double dz = rng.uniform(), dx = rng.uniform();
      Vec center(a + 0.9 * dx, 0.2, b + 0.9 * dz);
      Vec d = center - Vec(4, 0.2, 0);
      if (!(std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z) > 0.9))
        continue;
      if (choose_mat < 0.8) {
        Vec c2 = rng.vec(), c1 = rng.vec(); // albedo = random() * random(), right operand drawn first
        Vec albedo(c1.x * c2.x, c1.y * c2.y, c1.z * c2.z);
        Vec center2 = center + Vec(0, rng.uniform(0, .5), 0);
        int m;
        if (textured && rng.uniform() < 0.25)
          m = (textured_count++ % 2 == 0) ? marble : check;
        else
          m = s.lambertian(albedo);
        s.moving_sphere(center, center2, 0.2, m);
      } else if (choose_mat < 0.95) {
        Vec albedo = rng.vec(0.5, 1);
        double fuzz = rng.uniform(0, 0.5);
        s.sphere(center, 0.2, s.metal(albedo, fuzz));
      } else {
        s.sphere(center, 0.2, s.dielectric(1.5));
      }
    }
  }

  s.sphere(Vec(0, 1, 0), 1.0, s.dielectric(1.5));
  s.sphere(Vec(-4, 1, 0), 1.0, s.lambertian(Vec(0.4, 0.2, 0.1)));
  s.sphere(Vec(4, 1, 0), 1.0, s.metal(Vec(0.7, 0.6, 0.5), 0.0));

  set_camera(s, 16.0 / 9.0, Vec(0.70, 0.80, 1.00), 20, Vec(13, 2, 3), Vec(0, 0, 0), 0.6, 10.0);
}

void cornell_walls(SceneBuilder &s) {
  int red = s.lambertian(Vec(.65, .05, .05));
  int white = s.lambertian(Vec(.73, .73, .73));
  int green = s.lambertian(Vec(.12, .45, .15));
  int light = s.diffuse_light(Vec(15, 15, 15));
  s.quad(Vec(555, 0, 0), Vec(0, 0, 555), Vec(0, 555, 0), green);
  s.quad(Vec(0, 0, 555), Vec(0, 0, -555), Vec(0, 555, 0), red);
  s.quad(Vec(0, 555, 0), Vec(555, 0, 0), Vec(0, 0, 555), white);
  s.quad(Vec(0, 0, 555), Vec(555, 0, 0), Vec(0, 0, -555), white);
  s.quad(Vec(555, 0, 555), Vec(-555, 0, 0), Vec(0, 555, 0), white);
  s.quad(Vec(213, 554, 227), Vec(130, 0, 0), Vec(0, 0, 105), light);
  set_camera(s, 1.0, Vec(0, 0, 0), 40, Vec(278, 278, -800), Vec(278, 278, 0), 0, 10);
}

// populate_cornell_box_scene (main.cpp:21-71).
void cornell_scene(SceneBuilder &s) {
  cornell_walls(s);
  int white = s.lambertian(Vec(.73, .73, .73));
  s.box(Vec(0, 0, 0), Vec(165, 330, 165), white, {translate(Vec(265, 0, 295)), rotate_y(15)});
  s.sphere(Vec(190, 90, 190), 90, s.dielectric(1.5));
  s.light_quad(Vec(343, 554, 332), Vec(-130, 0, 0), Vec(0, 0, -105));
  s.light_sphere(Vec(190, 90, 190), 90);
}

// Config 3: the two blocks of the Cornell box as constant-density smoke.
void cornell_smoke_scene(SceneBuilder &s) {
  cornell_walls(s);
  s.box_medium(Vec(0, 0, 0), Vec(165, 330, 165), {translate(Vec(265, 0, 295)), rotate_y(15)}, 0.01, Vec(0, 0, 0));
  s.box_medium(Vec(0, 0, 0), Vec(165, 165, 165), {translate(Vec(130, 0, 65)), rotate_y(-18)}, 0.01, Vec(1, 1, 1));
  s.light_quad(Vec(343, 554, 332), Vec(-130, 0, 0), Vec(0, 0, -105));
}

// Config 5: ground of random-height boxes, area light, moving / glass / metal spheres, a glass sphere
// filled with blue medium, a global mist, a checker globe (the reference has no image texture), a
// marble sphere and a rotated, translated cluster of small spheres.
void final_scene(SceneBuilder &s, Rng &rng, int boxes_per_side, int n_cluster) {
  int ground = s.lambertian(Vec(0.48, 0.83, 0.53));
  for (int i = 0; i < boxes_per_side; i++)
    for (int j = 0; j < boxes_per_side; j++) {
      double w = 100.0;
      double x0 = -1000.0 + i * w, z0 = -1000.0 + j * w, y0 = 0.0;
      double x1 = x0 + w, y1 = rng.uniform(1, 101), z1 = z0 + w;
      s.box(Vec(x0, y0, z0), Vec(x1, y1, z1), ground, {});
    }
  s.quad(Vec(123, 554, 147), Vec(300, 0, 0), Vec(0, 0, 265), s.diffuse_light(Vec(7, 7, 7)));
  s.light_quad(Vec(123, 554, 147), Vec(300, 0, 0), Vec(0, 0, 265));

  Vec center1(400, 400, 200);
  s.moving_sphere(center1, center1 + Vec(30, 0, 0), 50, s.lambertian(Vec(0.7, 0.3, 0.1)));
  s.sphere(Vec(260, 150, 45), 50, s.dielectric(1.5));
  s.sphere(Vec(0, 150, 145), 50, s.metal(Vec(0.8, 0.8, 0.9), 1.0));

  s.sphere(Vec(360, 150, 145), 70, s.dielectric(1.5));
  s.sphere_medium(Vec(360, 150, 145), 70, 0.2, Vec(0.2, 0.4, 0.9));
  s.sphere_medium(Vec(0, 0, 0), 5000, 0.0001, Vec(1, 1, 1));

  s.sphere(Vec(400, 200, 400), 100, s.lambertian_tex(s.checker(20.0, Vec(.1, .2, .7), Vec(.9, .9, .9))));
  int table = s.perlin(rng);
  s.sphere(Vec(220, 280, 300), 80, s.lambertian_tex(s.noise(0.2, table)));

  int white = s.lambertian(Vec(.73, .73, .73));
  int xf = s.xform({translate(Vec(-100, 270, 395)), rotate_y(15)});
  for (int j = 0; j < n_cluster; j++)
    s.sphere(rng.vec(0, 165), 10, white, xf);

  set_camera(s, 16.0 / 9.0, Vec(0, 0, 0), 40, Vec(478, 278, -600), Vec(278, 278, 0), 0, 10);
}

// Image-texture scene ("The Next Week" earth()): a globe and a framed picture on a quad, both textured with
// one procedurally drawn map (the image has no network to fetch earthmap.jpg from): oceans, continents from
// a few overlapping discs in longitude / latitude, polar caps and a 15-degree graticule.  The reference has
// no image texture; this scene exists for the north star's image-texture requirement.
void earth_scene(SceneBuilder &s) {
  const int W = 512, H = 256;
  std::vector<uint8_t> rgb(size_t(W) * H * 3);
  const double blobs[][3] = {{-100, 45, 32}, {-60, -15, 24}, {20, 5, 28}, {25, 50, 22}, {90, 45, 38},
                             {135, -25, 16}, {-45, 72, 12}, {105, 15, 14}}; // lon, lat, radius (degrees)
  for (int j = 0; j < H; j++)
    for (int i = 0; i < W; i++) {
      double lon = (i + 0.5) / W * 360.0 - 180.0, lat = 90.0 - (j + 0.5) / H * 180.0;
      bool land = false;
      for (const auto &b : blobs) {
        double dl = std::fabs(lon - b[0]);
        dl = dl > 180 ? 360 - dl : dl;
        double dx = dl * std::cos(lat * 3.14159265358979323846 / 180.0), dy = lat - b[1];
        land = land || dx * dx + dy * dy < b[2] * b[2];
      }
      uint8_t r = 20, g = 60, bl = 150; // ocean
      if (land) {
        r = uint8_t(60 + std::fabs(lat) * 0.8);
        g = uint8_t(140 - std::fabs(lat) * 0.6);
        bl = 50;
      }
      if (std::fabs(lat) > 75) {
        r = g = bl = 235;
      }
      if (std::fmod(lon + 180.0, 15.0) < 0.5 || std::fmod(lat + 90.0, 15.0) < 0.5) {
        r = uint8_t(r / 2);
        g = uint8_t(g / 2);
        bl = uint8_t(bl / 2);
      }
      uint8_t *px = &rgb[(size_t(j) * W + i) * 3];
      px[0] = r;
      px[1] = g;
      px[2] = bl;
    }
  int map = s.image_texture(s.image(W, H, rgb));
  s.sphere(Vec(0, 0, 0), 2, s.lambertian_tex(map));
  s.quad(Vec(-6, -2.5, -4), Vec(5, 0, 0), Vec(0, 2.5, 0), s.lambertian_tex(map));
  s.sphere(Vec(0, -1002, 0), 1000, s.lambertian(Vec(0.5, 0.5, 0.5)));
  s.quad(Vec(3, 1, -3), Vec(2, 0, 0), Vec(0, 0, 2), s.diffuse_light(Vec(4, 4, 4)));
  s.light_quad(Vec(3, 1, -3), Vec(2, 0, 0), Vec(0, 0, 2));
  set_camera(s, 16.0 / 9.0, Vec(0.70, 0.80, 1.00), 30, Vec(0, 1, 12), Vec(0, 0, 0), 0, 10);
}

} // namespace

bool build_builtin(SceneBuilder &s, const std::string &name, uint64_t seed, int p0, int p1) {
  Rng rng(seed);
  if (name == "spheres")
    spheres_scene(s, rng, p0 > 0 ? p0 : 11, false);
  else if (name == "spheres_textured")
    spheres_scene(s, rng, p0 > 0 ? p0 : 11, true);
  else if (name == "cornell")
    cornell_scene(s);
  else if (name == "cornell_smoke")
    cornell_smoke_scene(s);
  else if (name == "final")
    final_scene(s, rng, p0 > 0 ? p0 : 20, p1 >= 0 ? p1 : 1000);
  else if (name == "earth")
    earth_scene(s);
  else
    return false;
  s.finalize();
  return true;
}

} // namespace rth
