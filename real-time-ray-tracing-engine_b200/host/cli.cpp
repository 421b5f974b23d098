// Command line of the `raytracer` program: the reference's flags with the reference's defaults
// (input/CLI.cpp:4-126, input/CLI.hpp:8-27) plus --scene / --seed / --gpus / --frames / --keys / --headless / --adaptive.
#include "../../include/rt_host.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace rth {
void set_error(const std::string &msg);
}

extern "C" {

const char *rth_cli_help(void) {
  return "Raytracer: B200-native path-tracing backend.\n"
         "Renders scenes using either a static or dynamic camera.\n\n"
         "Usage: raytracer [options]\n\n"
         "Options:\n"
         "  -h, --help                 Show this help message\n"
         "  --camera [static|dynamic]  Select camera type (default: static)\n"
         "  --output <file>            Output file name for static camera (default: image.ppm)\n"
         "  -p, --parallel             Accepted for compatibility (the GPU backend is always parallel)\n"
         "  -b, --bvh                  Accepted for compatibility (the GPU backend always uses its BVH)\n"
         "  -g, --gpu                  Accepted for compatibility (there is no CPU backend)\n"
         "  -d, --debug                Debug mode: dumps the scene as JSON to logs/scene_debug.json\n"
         "  --width <int>              Image width (default: 600)\n"
         "  --samples <int>            Samples per pixel (default: 100)\n"
         "  --depth <int>              Maximum ray bounces (default: 50)\n"
         "  --scene <name|file.json>   Built-in scene (spheres, spheres_textured, cornell, cornell_smoke, final)\n"
         "                             or a JSON scene file (default: cornell)\n"
         "  --seed <int>               Seed of the scene generator and of the render (default: 1234)\n"
         "  --gpus <int>               GPUs to partition the image over (default: 1)\n"
         "  --frames <int>             Dynamic camera without a window: progressive frames to render\n"
         "                             (default: one per stratum)\n"
         "  --headless                 Dynamic camera: never open a window (default: an SDL3 window when SDL3 is\n"
         "                             installed and --frames is not given)\n"
         "  --adaptive                 Headless frames adapt the samples per frame to the frame rate as the\n"
         "                             window does (> 30 FPS: twice as many, < 15 FPS: half, 1..64)\n"
         "  --keys <string>            Dynamic camera without a window: scripted key states, one character per\n"
         "                             frame (w/s/a/d move the camera by 10 units and restart the accumulation,\n"
         "                             +/- change the samples per pixel, anything else: no key)\n\n"
         "Examples:\n"
         "  raytracer --camera static --output render.ppm --scene spheres --width 400\n"
         "  raytracer --camera dynamic --scene spheres --width 1920 --samples 16 --depth 8\n";
}

int rth_cli_parse(int argc, char **argv, rth_cli_options *out) {
  if (!out)
    return 1;
  std::memset(out, 0, sizeof *out);
  out->width = 600;
  out->samples = 100;
  out->depth = 50;
  out->seed = 1234;
  out->gpus = 1;
  std::snprintf(out->output, sizeof out->output, "image.ppm");
  std::snprintf(out->scene, sizeof out->scene, "cornell");
  std::string errors;
  auto need_int = [&](int &i, const char *flag, int &dst) {
    if (i + 1 >= argc) {
      errors += std::string(flag) + " requires a number\n";
      return;
    }
    try {
      dst = std::stoi(argv[++i]);
    } catch (...) {
      errors += std::string(flag) + " requires a valid integer\n";
    }
  };
  auto need_str = [&](int &i, const char *flag, char *dst, size_t cap) {
    if (i + 1 >= argc) {
      errors += std::string(flag) + " requires an argument\n";
      return;
    }
    std::snprintf(dst, cap, "%s", argv[++i]);
  };
  for (int i = 1; i < argc; ++i) {
    std::string arg = argv[i];
    if (arg == "-h" || arg == "--help") {
      out->help = 1;
    } else if (arg == "--camera") {
      if (i + 1 < argc) {
        std::string type = argv[++i];
        if (type == "static")
          out->camera_dynamic = 0;
        else if (type == "dynamic")
          out->camera_dynamic = 1;
        else
          errors += "Unknown camera type: " + type + "\n";
      } else {
        errors += "--camera requires an argument: static or dynamic\n";
      }
    } else if (arg == "--output") {
      need_str(i, "--output", out->output, sizeof out->output);
    } else if (arg == "-p" || arg == "--parallel") {
      out->use_parallelism = 1;
    } else if (arg == "-b" || arg == "--bvh") {
      out->use_bvh = 1;
    } else if (arg == "-g" || arg == "--gpu") {
      out->use_gpu = 1;
    } else if (arg == "-d" || arg == "--debug") {
      out->debug = 1;
    } else if (arg == "--width") {
      need_int(i, "--width", out->width);
    } else if (arg == "--samples") {
      need_int(i, "--samples", out->samples);
    } else if (arg == "--depth") {
      need_int(i, "--depth", out->depth);
    } else if (arg == "--scene") {
      need_str(i, "--scene", out->scene, sizeof out->scene);
    } else if (arg == "--seed") {
      int v = 1234;
      need_int(i, "--seed", v);
      out->seed = (uint64_t)v;
    } else if (arg == "--gpus") {
      need_int(i, "--gpus", out->gpus);
    } else if (arg == "--frames") {
      need_int(i, "--frames", out->frames);
    } else if (arg == "--keys") {
      need_str(i, "--keys", out->keys, sizeof out->keys);
    } else if (arg == "--headless") {
      out->headless = 1;
    } else if (arg == "--adaptive") {
      out->adaptive = 1;
    } else {
      errors += "Unknown option: " + arg + "\n";
    }
  }
  if (out->width < 1 || out->samples < 1 || out->depth < 1 || out->gpus < 1)
    errors += "--width, --samples, --depth and --gpus must be positive\n";
  if (!errors.empty()) {
    rth::set_error(errors);
    return 1;
  }
  return 0;
}

} // extern "C"
