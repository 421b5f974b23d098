// TEST BUILD ONLY: the device functions of csrc/ (through emu.cpp) and the host scene code, compiled with
// -fsanitize=address,undefined and run over every built-in scene (tests/test_sanitizers_host.py).
// compute-sanitizer is closed on the GPU pool, so memory safety of the per-thread device code is checked here.
#include "emu.cpp"
#include "../../include/rt_host.h"
#include <cstdio>
extern "C" void ora_camera_init(const rt_camera_config *, rt_camera *);
int main(int argc, char **argv) {
  const std::string tmp = argc > 1 ? argv[1] : "/tmp";
  const char *names[] = {"spheres", "spheres_textured", "cornell", "cornell_smoke", "final", "earth"};
  int p0s[] = {11, 30, 0, 0, 6, 0};
  for (int k = 0; k < 6; k++) {
    rth_scene *hs = rth_scene_builtin(names[k], 1234, p0s[k], k == 4 ? 100 : -1);
    { // JSON writer and reader under the sanitizers: save, load back, compare the counts
      std::string path = tmp + "/asan_" + names[k] + ".json";
      if (rth_scene_save_json(hs, path.c_str()) != 0) { printf("%s: save failed %s\n", names[k], rth_last_error()); return 1; }
      rth_scene *back = rth_scene_load_json(path.c_str());
      if (!back) { printf("%s: load failed %s\n", names[k], rth_last_error()); return 1; }
      const rt_scene_desc *a = rth_scene_desc(hs), *b = rth_scene_desc(back);
      if (a->n_spheres != b->n_spheres || a->n_quads != b->n_quads || a->n_media != b->n_media || a->n_materials != b->n_materials) {
        printf("%s: JSON round trip changed the scene\n", names[k]);
        return 1;
      }
      rth_scene_free(back);
    }
    const rt_scene_desc *d = rth_scene_desc(hs);
    EmuScene *es = emu_scene_create(d);
    if (!es) { printf("%s: create failed %s\n", names[k], emu_last_error()); return 1; }
    rt_camera_config cfg; rth_scene_camera(hs, 96, 4, 12, &cfg);
    rt_camera cam; 
    ora_camera_init(&cfg, &cam);
    std::vector<float> out((size_t)cam.image_width * cam.image_height * 3);
    uint64_t segs = 0;
    emu_render(es, &cam, 0, 4, 2, 12, 3, 0.25, out.data(), &segs);
    double m = 0; for (float v : out) m += v;
    // garbage rays (NaN / infinite / zero components): the FP32 traversal must stay inside the arrays - an unused
    // child slot is unreachable for any ray (RT_EMPTY is +inf as a float, rt_device.h)
    {
      const float nan = __builtin_nanf(""), inf = __builtin_huge_valf();
      const float vals[] = {0.f, -0.f, 1.f, -1.f, nan, inf, -inf, 1e30f, 1e-30f};
      std::vector<rt_ray> rays;
      for (float a : vals)
        for (float b : vals)
          for (float c : vals) {
            rt_ray r{};
            r.origin[0] = b, r.origin[1] = 1.0, r.origin[2] = c;
            r.direction[0] = a, r.direction[1] = c, r.direction[2] = b;
            r.time = a == a ? 0.5 : a;
            r.t_min = 0.001, r.t_max = inf;
            rays.push_back(r);
            r.origin[1] = a;
            rays.push_back(r);
          }
      std::vector<rt_hit> hits(rays.size());
      emu_trace(es, rays.data(), (int64_t)rays.size(), RT_TRACE_FAST_F32, 1, hits.data());
      for (const rt_hit &h : hits)
        if (h.prim < -1 || h.prim >= d->n_spheres + d->n_quads + d->n_media) { printf("%s: bad hit %d\n", names[k], h.prim); return 1; }
    }
    // animated scene: move the first surface spheres, refit, render again (rt_scene_update_spheres)
    if (d->n_spheres > 4 && !(d->spheres[1].flags & RT_PRIM_BOUNDARY) && !(d->spheres[2].flags & RT_PRIM_BOUNDARY)) {
      rt_sphere moved[2] = {d->spheres[1], d->spheres[2]};
      moved[0].center0[0] += 1.5, moved[1].center0[2] -= 2.0, moved[1].radius *= 1.5;
      if (emu_scene_update_spheres(es, 1, 2, moved) != 0 || emu_scene_check_bvh(es) != 0) { printf("%s: refit failed\n", names[k]); return 1; }
      emu_render(es, &cam, 0, 1, 2, 6, 3, 1.0, out.data(), &segs);
    }
    printf("%s: bvh check %d, segs %llu, mean %.4f\n", names[k], emu_scene_check_bvh(es), (unsigned long long)segs, m / out.size());
    emu_scene_destroy(es); rth_scene_free(hs);
  }
  return 0;
}
