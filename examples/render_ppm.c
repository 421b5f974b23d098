/* examples/render_ppm.c - the C ABI from plain C11: build a scene with the host library, render it with
 * librt_b200.so the way StaticCamera::render_gpu drives the reference's CUDA backend
 * (core/camera/StaticCamera.cpp:160-301), write the reference's ASCII PPM.
 *
 *   gcc -std=c11 -Wall -Wextra -pedantic -Iinclude examples/render_ppm.c \
 *       -Lreal-time-ray-tracing-engine_b200/host -lrt_host -Lreal-time-ray-tracing-engine_b200/csrc -lrt_b200 -lm -o render_ppm
 *   ./render_ppm cornell 300 64 20 out.ppm
 *
 * Exit status 2 when there is no CUDA device: this backend has no CPU rendering path. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rt_b200.h"
#include "rt_host.h"

#define CHECK(call)                                                                   \
  do {                                                                                \
    int st_ = (call);                                                                 \
    if (st_ != RT_OK) {                                                               \
      fprintf(stderr, "[ERROR] %s failed (%d): %s\n", #call, st_, rt_last_error());   \
      return st_ == RT_ERR_NO_DEVICE ? 2 : 1;                                         \
    }                                                                                 \
  } while (0)

int main(int argc, char **argv) {
  const char *scene_name = argc > 1 ? argv[1] : "cornell";
  int width = argc > 2 ? atoi(argv[2]) : 300;
  int samples = argc > 3 ? atoi(argv[3]) : 64;
  int depth = argc > 4 ? atoi(argv[4]) : 20;
  const char *out = argc > 5 ? argv[5] : "image.ppm";

  rth_scene *hs = rth_scene_builtin(scene_name, 1234, 0, -1);
  if (!hs) {
    fprintf(stderr, "[ERROR] %s\n", rth_last_error());
    return 1;
  }
  rt_camera_config cfg;
  rth_scene_camera(hs, width, samples, depth, &cfg);
  rt_camera cam;
  CHECK(rt_camera_init(&cfg, &cam)); /* Camera::initialize */

  rt_context *ctx = NULL;
  rt_scene *scene = NULL;
  rt_film *film = NULL;
  CHECK(rt_context_create(0, &ctx));
  CHECK(rt_scene_create(ctx, rth_scene_desc(hs), &scene));                                  /* initialize_cuda_scene */
  CHECK(rt_film_create(ctx, cam.image_width, cam.image_height, 0, 1, 8, NULL, &film));
  int sqrt_spp = (int)sqrt((double)samples);
  CHECK(rt_render_static(scene, &cam, film, sqrt_spp, depth, 1234));                        /* cuda_static_render_wrapper */
  unsigned char *rgb8 = malloc((size_t)cam.image_width * (size_t)cam.image_height * 3);
  if (!rgb8)
    return 1;
  CHECK(rt_film_resolve_rgb8(film, 1.0 / (sqrt_spp * sqrt_spp), rgb8));                     /* write_color's to_byte */
  if (rth_write_ppm_p3(out, cam.image_width, cam.image_height, rgb8) != 0) {
    fprintf(stderr, "[ERROR] %s\n", rth_last_error());
    return 1;
  }
  printf("%s: %dx%d, %d spp, depth %d -> %s\n", scene_name, cam.image_width, cam.image_height, sqrt_spp * sqrt_spp, depth, out);
  free(rgb8);
  rt_film_destroy(film);
  rt_scene_destroy(scene);
  rt_context_destroy(ctx);
  rth_scene_free(hs);
  return 0;
}
