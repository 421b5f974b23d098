// TEST BUILD ONLY: the device functions of csrc/ (through emu.cpp) and the host scene code, compiled with
// -fsanitize=address,undefined and run over every built-in scene (tests/test_sanitizers_host.py).
// compute-sanitizer is closed on the GPU pool, so memory safety of the per-thread device code is checked here.
#include "emu.cpp"
#include "../../include/rt_host.h"
#include <cstdio>
extern "C" void ora_camera_init(const rt_camera_config *, rt_camera *);
int main() {
  const char *names[] = {"spheres", "spheres_textured", "cornell", "cornell_smoke", "final", "earth"};
  int p0s[] = {11, 30, 0, 0, 6, 0};
  for (int k = 0; k < 6; k++) {
    rth_scene *hs = rth_scene_builtin(names[k], 1234, p0s[k], k == 4 ? 100 : -1);
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
    printf("%s: bvh check %d, segs %llu, mean %.4f\n", names[k], emu_scene_check_bvh(es), (unsigned long long)segs, m / out.size());
    emu_scene_destroy(es); rth_scene_free(hs);
  }
  return 0;
}
