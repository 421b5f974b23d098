/*
 * rt_host.h — C interface of the host-side library (librt_host.so): scene construction, the JSON
 * scene format, the reference's CLI options and its PPM output.  Everything here is plain host C++
 * behind a C ABI; it produces the flat rt_scene_desc that librt_b200.so (rt_b200.h) consumes and
 * never touches the GPU.
 *
 * Reference interfaces mirrored (paths relative to the reference's src/):
 *   rth_scene_builtin     <- populate_cornell_box_scene / populate_bouncing_spheres_scene
 *                            (main.cpp:21-131): the reference's scenes are hard-coded C++ functions
 *   rth_scene_load_json   <- the JSON scene format the reference's README advertises; the reference
 *   rth_scene_save_json      itself only *dumps* JSON (Camera.cpp:75-150), so the schema mirrors the
 *                            dump's class and field names
 *   rth_cli_parse         <- parse_cli / CLIOptions (input/CLI.cpp:4-92, input/CLI.hpp:8-51)
 *   rth_write_ppm_p3      <- the PPM P3 writer in StaticCamera::render_cpu (StaticCamera.cpp:57,94-99,
 *                            utils/ColorUtility.hpp:30-37)
 *   rth_presenter_*       <- the SDL3 window of DynamicCamera (core/camera/DynamicCamera.cpp:62-91,196-306)
 */
#ifndef RT_HOST_H
#define RT_HOST_H

#include "rt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rth_scene rth_scene;

const char *rth_last_error(void);

/* Built-in scenes, generated with std::mt19937(seed) drawing in the reference's order:
 *   "spheres"           main.cpp:73-131; p0 = grid half size (<= 0: the reference's 11)
 *   "spheres_textured"  the same generator, a quarter of the diffuse spheres marble / checker (config 4)
 *   "cornell"           main.cpp:21-71
 *   "cornell_smoke"     Cornell box with two constant-density media (config 3)
 *   "final"             boxes + light + media + textured spheres + sphere cluster (config 5);
 *                       p0 = boxes per side (<= 0: 20), p1 = cluster spheres (< 0: 1000)
 *   "earth"             a globe and a picture quad with an image texture (not in the reference)
 * Returns NULL (and sets rth_last_error) for an unknown name. */
rth_scene *rth_scene_builtin(const char *name, uint64_t seed, int p0, int p1);
rth_scene *rth_scene_load_json(const char *path);
int rth_scene_save_json(const rth_scene *scene, const char *path);
void rth_scene_free(rth_scene *scene);

const rt_scene_desc *rth_scene_desc(const rth_scene *scene);
/* The scene's camera with the CLI-controlled fields filled in (main.cpp:151-159). */
void rth_scene_camera(const rth_scene *scene, int image_width, int samples_per_pixel, int max_depth,
                      rt_camera_config *out);

/* CLIOptions (input/CLI.hpp:8-27) plus the additions of this build. */
typedef struct rth_cli_options {
  int width;          /* --width, default 600 */
  int samples;        /* --samples, default 100 */
  int depth;          /* --depth, default 50 */
  int camera_dynamic; /* --camera static|dynamic, default static */
  int use_parallelism;/* -p / --parallel (accepted; the GPU backend is always parallel) */
  int use_bvh;        /* -b / --bvh (accepted; the GPU backend always uses its BVH) */
  int use_gpu;        /* -g / --gpu (accepted; there is no CPU backend) */
  int debug;          /* -d / --debug */
  int help;           /* -h / --help */
  char output[256];   /* --output, default "image.ppm" */
  /* additions */
  char scene[256];    /* --scene <builtin name or file.json>, default "cornell" (main.cpp:161) */
  uint64_t seed;      /* --seed, default 1234 */
  int gpus;           /* --gpus, default 1 */
  int frames;         /* --frames: dynamic mode without a window renders this many frames, default 0 = sqrt_spp^2 */
  char keys[256];     /* --keys: scripted key states for the headless dynamic camera, one character per frame:
                         w/s/a/d move lookfrom and lookat by 10 units along +z/-z/-x/+x and restart the
                         accumulation, '+'/'-' change samples per pixel (DynamicCamera::handle_events,
                         core/camera/DynamicCamera.cpp:204-278), any other character = no key */
  int headless;       /* --headless: the dynamic camera never opens a window (default: a window when SDL3 can be loaded
                         and --frames is not given) */
  int adaptive;       /* --adaptive: headless frames adapt the samples per frame to the frame rate as the window does */
} rth_cli_options;

/* Returns 0 on success, non-zero on a malformed command line (message in rth_last_error). */
int rth_cli_parse(int argc, char **argv, rth_cli_options *out);
const char *rth_cli_help(void);

/* Progressive window of the dynamic camera (DynamicCamera.cpp:62-91 window + streaming RGB24 texture, :204-278 key
 * handling, :196-200 present).  SDL3 is opened with dlopen at run time (RT_SDL3_LIB overrides the library name):
 * rth_presenter_open returns NULL and sets rth_last_error when it is not installed. */
typedef struct rth_presenter rth_presenter;
typedef struct rth_input {
  int quit;      /* ESC pressed or window closed */
  int spp_delta; /* '=' presses minus '-' presses since the last poll */
  int move_x;    /* D held (+1) / A held (-1) */
  int move_z;    /* W held (+1) / S held (-1) */
  int moved;     /* any movement key held: the accumulation restarts (even if the moves cancel) */
} rth_input;
rth_presenter *rth_presenter_open(int width, int height, const char *title);
int rth_presenter_poll(rth_presenter *presenter, rth_input *out);
/* Uploads one row-major RGB8 frame and shows it; status_line (may be NULL) goes to the window title. */
int rth_presenter_present(rth_presenter *presenter, const uint8_t *rgb8, const char *status_line);
void rth_presenter_close(rth_presenter *presenter);

/* "P3\nW H\n255\n" followed by one "r g b\n" line per pixel. */
int rth_write_ppm_p3(const char *path, int width, int height, const uint8_t *rgb8);

#ifdef __cplusplus
}
#endif
#endif /* RT_HOST_H */
