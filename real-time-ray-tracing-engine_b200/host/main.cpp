// raytracer — the host program: the reference's CLI (src/main.cpp:133-173) in front of the B200
// backend.  Static camera: renders the image and writes output/<file> as ASCII PPM
// (StaticCamera::render, core/camera/StaticCamera.cpp:25-57,94-99).  Dynamic camera: the reference
// opens an SDL3 window and adds one stratum per frame (DynamicCamera.cpp:103-194).  Here the window is opened
// when libSDL3 can be loaded at run time (host/presenter.cpp, dlopen) and --frames is not given; otherwise the
// dynamic camera runs headless: it renders the progressive frames, resolves each to RGB8 exactly as
// update_texture does (:280-306), takes its key states from --keys, reports per-frame times and writes the
// last frame.  Images are partitioned over --gpus devices by interleaved scanline tiles; every GPU tone-maps its
// tiles and stores them straight into GPU 0's frame over NVLink (rt_film_present), and the frame's copy to pinned
// host memory runs beside the next frame's render.  There is no CPU rendering path: without a CUDA device the
// program fails.
#include "../../include/rt_b200.h"
#include "../../include/rt_host.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <sys/stat.h>
#include <vector>

#define CHECK(call)                                                                                          \
  do {                                                                                                       \
    int st_ = (call);                                                                                        \
    if (st_ != RT_OK) {                                                                                      \
      std::fprintf(stderr, "[ERROR] %s failed (%d): %s\n", #call, st_, rt_last_error());                      \
      return 1;                                                                                              \
    }                                                                                                        \
  } while (0)

static double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
  rth_cli_options opt;
  if (rth_cli_parse(argc, argv, &opt) != 0) {
    std::fprintf(stderr, "%s", rth_last_error());
    std::fputs(rth_cli_help(), stdout);
    return 1;
  }
  if (opt.help) {
    std::fputs(rth_cli_help(), stdout);
    return 0;
  }
  std::string scene_name = opt.scene;
  bool is_file = scene_name.size() > 5 && scene_name.substr(scene_name.size() - 5) == ".json";
  rth_scene *hs = is_file ? rth_scene_load_json(scene_name.c_str()) : rth_scene_builtin(scene_name.c_str(), opt.seed, 0, -1);
  if (!hs) {
    std::fprintf(stderr, "[ERROR] %s\n", rth_last_error());
    return 1;
  }
  if (opt.debug) {
    mkdir("logs", 0755);
    rth_scene_save_json(hs, "logs/scene_debug.json");
    std::fprintf(stderr, "[DEBUG] scene written to logs/scene_debug.json\n");
  }
  rt_camera_config cfg;
  rth_scene_camera(hs, opt.width, opt.samples, opt.depth, &cfg);
  rt_camera cam;
  CHECK(rt_camera_init(&cfg, &cam));
  const int W = cam.image_width, H = cam.image_height;
  int sqrt_spp = int(std::sqrt(double(opt.samples))); // Camera.cpp:209
  const int n_gpus = opt.gpus;
  if (rt_device_count() < n_gpus) {
    std::fprintf(stderr, "[ERROR] %d CUDA device(s) requested, %d available; there is no CPU fallback\n", n_gpus,
                 rt_device_count());
    return 1;
  }

  std::vector<rt_context *> ctx(n_gpus);
  std::vector<rt_scene *> scene(n_gpus);
  std::vector<rt_film *> film(n_gpus);
  const int tile_rows = 8;
  for (int g = 0; g < n_gpus; g++) {
    CHECK(rt_context_create(g, &ctx[g]));
    if (opt.camera_dynamic) // one graph launch per progressive frame instead of one launch per kernel
      CHECK(rt_context_set_graph(ctx[g], 1));
    CHECK(rt_scene_create(ctx[g], rth_scene_desc(hs), &scene[g])); // the scene is replicated on every GPU
    CHECK(rt_film_create(ctx[g], W, H, g, n_gpus, tile_rows, nullptr, &film[g]));
  }
  rt_scene_info info;
  CHECK(rt_scene_get_info(scene[0], &info));
  std::fprintf(stderr, "[INFO] scene %s: %lld primitives, %lld BVH4 nodes, device build %.2f ms; image %dx%d, %d spp, depth %d, %d GPU(s)\n",
               opt.scene, (long long)info.n_prims, (long long)info.n_nodes, info.build_ms, W, H, sqrt_spp * sqrt_spp,
               opt.depth, n_gpus);

  // The displayed frame: two RGB8 frames in GPU 0's memory (the copy of one to host memory overlaps the render of
  // the next), every other GPU holds a handle to them and stores its tiles there over NVLink (rt_film_present);
  // two pinned host buffers receive them.  Nothing is allocated per frame.
  rt_frame *frame[2] = {nullptr, nullptr};
  std::vector<rt_frame *> handle[2] = {std::vector<rt_frame *>(n_gpus, nullptr), std::vector<rt_frame *>(n_gpus, nullptr)};
  uint8_t *rgb8_buf[2] = {nullptr, nullptr};
  for (int k = 0; k < 2; k++) {
    CHECK(rt_frame_create(ctx[0], W, H, n_gpus, &frame[k]));
    handle[k][0] = frame[k];
    for (int g = 1; g < n_gpus; g++)
      CHECK(rt_frame_attach(ctx[g], frame[k], &handle[k][g]));
    void *p = nullptr;
    CHECK(rt_host_alloc((size_t)W * H * 3, &p));
    rgb8_buf[k] = static_cast<uint8_t *>(p);
  }
  // every GPU tone-maps and stores its tiles, GPU 0 waits for all of them and starts the copy to host memory
  auto present = [&](int k, double scale) -> int {
    for (int g = 0; g < n_gpus; g++) {
      int st = rt_film_present(film[g], scale, handle[k][g]);
      if (st != RT_OK)
        return st;
    }
    int st = rt_frame_wait(frame[k]);
    return st != RT_OK ? st : rt_frame_download(frame[k], rgb8_buf[k]);
  };
  uint8_t *rgb8 = rgb8_buf[0]; // the most recent frame that is complete in host memory

  double t0 = now_ms();
  if (!opt.camera_dynamic) {
    for (int g = 0; g < n_gpus; g++)
      CHECK(rt_render_static(scene[g], &cam, film[g], sqrt_spp, opt.depth, opt.seed));
    CHECK(present(0, 1.0 / opt.samples)); // pixel_samples_scale = 1/spp
    CHECK(rt_frame_download_wait(frame[0]));
    double t1 = now_ms();
    long long paths = (long long)W * H * sqrt_spp * sqrt_spp;
    std::fprintf(stderr, "[INFO] rendered %lld path samples in %.1f ms (%.1f Mpath-samples/s)\n", paths, t1 - t0,
                 paths / (t1 - t0) / 1e3);
    mkdir("output", 0755);
    std::string path = std::string("output/") + opt.output;
    if (rth_write_ppm_p3(path.c_str(), W, H, rgb8) != 0) {
      std::fprintf(stderr, "[ERROR] %s\n", rth_last_error());
      return 1;
    }
    std::fprintf(stderr, "[INFO] wrote %s\n", path.c_str());
  } else {
    // Dynamic camera (DynamicCamera.cpp:103-200).  One displayed frame = handle input, add strata to the
    // accumulation unless it has converged, resolve to RGB8, present.  With a window the key states come from
    // SDL3 and the loop runs until ESC; headless they come from --keys and the loop runs --frames frames.
    rth_presenter *window = nullptr;
    if (!opt.headless && opt.frames <= 0) {
      window = rth_presenter_open(W, H, "Dynamic Camera");
      if (!window)
        std::fprintf(stderr, "[WARN] %s; rendering headless\n", rth_last_error());
    }
    const bool adaptive = window != nullptr || opt.adaptive;
    int frames = window ? -1 : (opt.frames > 0 ? opt.frames : sqrt_spp * sqrt_spp);
    int taken = 0, samples = opt.samples, moves = 0, shown = 0, last_queued = -1;
    // Adaptive quality: the reference doubles / halves its tile size (16..64 px) once a second when the frame
    // rate is above 30 / below 15 FPS (DynamicCamera.cpp:180-193).  A frame is one wavefront pass here, so the
    // knob is the number of strata added per displayed frame (1..64) with the same thresholds.
    int strata_per_frame = 1, fps_frames = 0;
    double fps = 0.0, fps_t0 = now_ms();
    const size_t n_keys = std::strlen(opt.keys);
    for (int f = 0; frames < 0 || f < frames; f++) {
      double f0 = now_ms();
      rth_input in{};
      if (window) {
        rth_presenter_poll(window, &in);
        if (in.quit)
          break;
      } else { // scripted key state of this frame
        char key = size_t(f) < n_keys ? opt.keys[f] : '.';
        in.move_x = key == 'd' ? 1 : (key == 'a' ? -1 : 0);
        in.move_z = key == 'w' ? 1 : (key == 's' ? -1 : 0);
        in.moved = in.move_x != 0 || in.move_z != 0;
        in.spp_delta = key == '+' ? 1 : (key == '-' ? -1 : 0);
      }
      // DynamicCamera::handle_events (:204-278)
      samples = std::max(1, samples + in.spp_delta);
      sqrt_spp = std::max(1, int(std::sqrt(double(samples))));
      const int total_strata = sqrt_spp * sqrt_spp;
      if (in.moved) { // camera moved: restart the accumulation and recompute the camera
        const double step = 10.0;
        cfg.lookfrom[0] += step * in.move_x, cfg.lookat[0] += step * in.move_x;
        cfg.lookfrom[2] += step * in.move_z, cfg.lookat[2] += step * in.move_z;
        CHECK(rt_camera_init(&cfg, &cam));
        for (int g = 0; g < n_gpus; g++)
          CHECK(rt_film_clear(film[g]));
        taken = 0;
        moves++;
      }
      // the window stops sampling at convergence (:115); headless frames keep refining with a new seed per round
      const bool converged = window != nullptr && taken >= total_strata;
      if (!converged) {
        int s = taken % total_strata;
        int n = adaptive ? std::min(strata_per_frame, total_strata - s) : 1;
        uint64_t seed = opt.seed + (uint64_t)(taken / total_strata);
        for (int g = 0; g < n_gpus; g++) {
          if (n == 1)
            CHECK(rt_render_accumulate(scene[g], &cam, film[g], s % sqrt_spp, s / sqrt_spp, sqrt_spp, opt.depth, seed));
          else
            CHECK(rt_render_strata(scene[g], &cam, film[g], s, n, sqrt_spp, opt.depth, seed));
        }
        taken += n;
      }
      double scale = 1.0 / std::max(1, taken); // DynamicCamera.cpp:285
      // Frame f is queued (render, tone map, NVLink stores, copy to pinned host memory) while frame f - 1, whose
      // copy ran beside this frame's render, is what gets shown: the host never waits for the GPU to go idle.
      const int k = f & 1;
      CHECK(present(k, scale));
      const bool have_previous = last_queued >= 0;
      if (have_previous) {
        CHECK(rt_frame_download_wait(frame[last_queued]));
        rgb8 = rgb8_buf[last_queued];
      }
      last_queued = k;
      shown++;
      fps_frames++;
      double f1 = now_ms();
      if (f1 - fps_t0 >= 1000.0) { // once a second (:180-193)
        fps = 1000.0 * fps_frames / (f1 - fps_t0);
        fps_frames = 0;
        fps_t0 = f1;
        if (adaptive && !converged && fps > 30.0 && strata_per_frame < 64)
          strata_per_frame *= 2;
        else if (adaptive && !converged && fps < 15.0 && strata_per_frame > 1)
          strata_per_frame /= 2;
      }
      if (window && have_previous) {
        char line[160]; // draw_fps (:308-348) as the window title
        std::snprintf(line, sizeof line, "Dynamic Camera - %.1f fps, %d/%d samples%s", fps, std::min(taken, total_strata),
                      total_strata, converged ? " - converged" : "");
        if (rth_presenter_present(window, rgb8, line) != 0) {
          std::fprintf(stderr, "[ERROR] %s\n", rth_last_error());
          break;
        }
      } else if (f < 5 || f == frames - 1) {
        std::fprintf(stderr, "[INFO] frame %d: %.3f ms (%.1f FPS)\n", f, f1 - f0, 1000.0 / (f1 - f0));
      }
    }
    if (last_queued >= 0) { // the last frame queued: shown (one frame behind, like all the others) and written out
      CHECK(rt_frame_download_wait(frame[last_queued]));
      rgb8 = rgb8_buf[last_queued];
      if (window)
        rth_presenter_present(window, rgb8, "Dynamic Camera");
    }
    if (window)
      rth_presenter_close(window);
    mkdir("output", 0755);
    std::string path = std::string("output/") + opt.output;
    rth_write_ppm_p3(path.c_str(), W, H, rgb8);
    std::fprintf(stderr, "[INFO] %d progressive frames in %.1f ms (%d camera move(s), %d sample(s) in the last "
                         "accumulation); last frame written to %s\n", shown, now_ms() - t0, moves, taken, path.c_str());
  }
  for (int k = 0; k < 2; k++) {
    for (int g = 1; g < n_gpus; g++)
      rt_frame_destroy(handle[k][g]);
    rt_frame_destroy(frame[k]);
    rt_host_free(rgb8_buf[k]);
  }
  for (int g = 0; g < n_gpus; g++) {
    rt_film_destroy(film[g]);
    rt_scene_destroy(scene[g]);
    rt_context_destroy(ctx[g]);
  }
  rth_scene_free(hs);
  return 0;
}
