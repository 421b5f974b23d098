// C interface of librt_host.so (include/rt_host.h): scenes, camera configuration, PPM output.
#include "../../include/rt_host.h"
#include "scene_builder.hpp"

#include <cstdio>
#include <cstring>
#include <string>

namespace rth {
thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
} // namespace rth

extern "C" {

const char *rth_last_error(void) { return rth::g_error.c_str(); }

rth_scene *rth_scene_builtin(const char *name, uint64_t seed, int p0, int p1) {
  rth_scene *s = new rth_scene();
  if (!name || !rth::build_builtin(s->builder, name, seed, p0, p1)) {
    rth::set_error(std::string("unknown built-in scene: ") + (name ? name : "(null)"));
    delete s;
    return nullptr;
  }
  return s;
}

void rth_scene_free(rth_scene *scene) { delete scene; }

const rt_scene_desc *rth_scene_desc(const rth_scene *scene) {
  return const_cast<rth_scene *>(scene)->builder.finalize();
}

void rth_scene_camera(const rth_scene *scene, int image_width, int samples_per_pixel, int max_depth,
                      rt_camera_config *out) {
  *out = scene->builder.camera;
  out->image_width = image_width;
  out->samples_per_pixel = samples_per_pixel;
  out->max_depth = max_depth;
  out->pad_ = 0;
}

// StaticCamera::render_cpu's output format (core/camera/StaticCamera.cpp:57,94-99;
// utils/ColorUtility.hpp:30-37): ASCII P3, one pixel per line.
int rth_write_ppm_p3(const char *path, int width, int height, const uint8_t *rgb8) {
  FILE *f = std::fopen(path, "w");
  if (!f) {
    rth::set_error(std::string("cannot open ") + path);
    return 1;
  }
  std::fprintf(f, "P3\n%d %d\n255\n", width, height);
  std::string buf;
  buf.reserve(size_t(width) * 12);
  for (int j = 0; j < height; j++) {
    buf.clear();
    for (int i = 0; i < width; i++) {
      const uint8_t *p = rgb8 + (size_t(j) * width + i) * 3;
      char line[16];
      int n = std::snprintf(line, sizeof line, "%d %d %d\n", int(p[0]), int(p[1]), int(p[2]));
      buf.append(line, size_t(n));
    }
    std::fwrite(buf.data(), 1, buf.size(), f);
  }
  std::fclose(f);
  return 0;
}

} // extern "C"
