/*
 * rt_b200.h — C ABI of the B200-native path-tracing backend (librt_b200.so).
 *
 * This is the drop-in boundary for the reference's GPU render seam.  The reference has no
 * plugin/FFI layer; its camera drivers call a handful of plain C++ functions.  Each entry point
 * below names the reference interface it replaces (paths relative to the reference's src/):
 *
 *   rt_scene_create / rt_scene_destroy   <- initialize_cuda_scene / cleanup_cuda_scene
 *                                           (scene/CudaSceneInitialization.cuh:249-308)
 *   (no equivalent needed)               <- cuda_init_rand_states_wrapper
 *                                           (core/camera/CameraKernelWrappers.cuh:11-13): the RNG is
 *                                           counter-based Philox, there is no per-pixel state to seed
 *   rt_render_accumulate                 <- cuda_dynamic_render_tile_wrapper
 *                                           (core/camera/CameraKernelWrappers.cuh:15-23), whole frame
 *                                           instead of one launch per tile
 *   rt_render_static                     <- cuda_static_render_wrapper
 *                                           (core/camera/CameraKernelWrappers.cuh:25-32), whole image
 *                                           instead of 64-row batches
 *   rt_film_resolve_rgb8                 <- DynamicCamera::update_texture (core/camera/DynamicCamera.cpp:280-306)
 *                                           and to_byte/linear_to_gamma (utils/ColorUtility.hpp:11-23)
 *   rt_camera_init                       <- Camera::initialize (core/camera/Camera.cpp:31-73)
 *   rt_trace_rays                        <- Hittable::hit on the world (core/Hittable.hpp:30-31),
 *                                           exposed as the closest-hit parity hook
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns an rt_status
 * (RT_OK = 0) and rt_last_error() returns a thread-local message for the last failure.  Handles are
 * opaque and thread-compatible (one context per GPU, calls on a handle from one thread at a time).
 * There is NO CPU fallback: without a usable CUDA device every call that needs one fails.
 *
 * All scene quantities are double precision (the reference's arithmetic type); the library derives
 * its own FP32 device layout from them.
 */
#ifndef RT_B200_H
#define RT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_B200_ABI_VERSION 2 /* 2: rt_counters.tail_segments + true node / primitive-test counts, parity audit, frames */

typedef enum rt_status {
  RT_OK = 0,
  RT_ERR_INVALID = 1,     /* bad argument / malformed scene */
  RT_ERR_CUDA = 2,        /* CUDA runtime error (see rt_last_error) */
  RT_ERR_NO_DEVICE = 3,   /* no CUDA device: there is no CPU fallback */
  RT_ERR_UNSUPPORTED = 4
} rt_status;

/* ----------------------------------------------------------------------------------------------
 * Scene description (flat arrays; the host side flattens the reference's object graph into it)
 * -------------------------------------------------------------------------------------------- */

enum { RT_MAT_LAMBERTIAN = 0, RT_MAT_METAL = 1, RT_MAT_DIELECTRIC = 2, RT_MAT_DIFFUSE_LIGHT = 3,
       RT_MAT_ISOTROPIC = 4 };
enum { RT_TEX_SOLID = 0, RT_TEX_CHECKER = 1, RT_TEX_NOISE = 2,
       RT_TEX_IMAGE = 3 /* not in the reference (it has no image texture); named by the north star */ };
enum { RT_XF_TRANSLATE = 0, RT_XF_ROTATE_Y = 1 };
enum { RT_SHAPE_SPHERE = 0, RT_SHAPE_QUAD = 1 };
/* rt_sphere.flags / rt_quad.flags */
enum { RT_PRIM_BOUNDARY = 1 }; /* primitive is (part of) a constant-medium boundary, not a surface */

#define RT_PERLIN_POINTS 256

/* Sphere (scene/objects/Sphere.hpp): centre moves linearly center0 -> center0 + center_dir over
 * the shutter interval [0,1) (center_dir = 0 for a static sphere). */
typedef struct rt_sphere {
  double center0[3];
  double center_dir[3];
  double radius;
  int32_t material; /* index into materials, -1 for boundary-only primitives */
  int32_t xform;    /* index into xforms, -1 = none */
  int32_t object;   /* index of the top-level world object this primitive belongs to */
  int32_t flags;
} rt_sphere;

/* Parallelogram "Plane" (scene/objects/Plane.hpp): corner + a*u + b*v, a,b in [0,1]. */
typedef struct rt_quad {
  double corner[3];
  double u[3];
  double v[3];
  int32_t material;
  int32_t xform;
  int32_t object;
  int32_t flags;
} rt_quad;

/* One instancing wrapper (scene/objects/Translate.hpp, RotateY.hpp). */
typedef struct rt_xform_op {
  int32_t type;     /* RT_XF_* */
  int32_t pad_;
  double offset[3]; /* translate */
  double angle_deg; /* rotate_y: the angle given to the RotateY constructor */
  double sin_theta; /* rotate_y: sin/cos of degrees_to_radians(angle) as RotateY.cpp:6-8 computes them */
  double cos_theta;
} rt_xform_op;

/* A chain of wrappers, outermost first: ops[first_op] is applied to the world-space ray first. */
typedef struct rt_xform {
  int32_t first_op;
  int32_t n_ops;
} rt_xform;

/* ConstantMedium (scene/mediums/ConstantMedium.hpp): a convex boundary made of n_prims primitives
 * of one shape type starting at first_prim in the spheres or quads array. */
typedef struct rt_medium {
  double density;
  int32_t shape;      /* RT_SHAPE_SPHERE or RT_SHAPE_QUAD */
  int32_t first_prim;
  int32_t n_prims;
  int32_t material;   /* phase function (RT_MAT_ISOTROPIC) */
  int32_t object;
  int32_t pad_;
} rt_medium;

typedef struct rt_material {
  int32_t type;    /* RT_MAT_* */
  int32_t texture; /* lambertian / diffuse_light / isotropic: index into textures */
  double albedo[3];/* metal */
  double fuzz;     /* metal */
  double ior;      /* dielectric refraction index */
} rt_material;

typedef struct rt_texture {
  int32_t type;    /* RT_TEX_* */
  int32_t even;    /* checker: texture indices */
  int32_t odd;
  int32_t perlin;  /* noise: index into perlins; image: index into images */
  double color[3]; /* solid */
  double scale;    /* checker / noise */
} rt_texture;

typedef struct rt_perlin {
  double rand_vec[RT_PERLIN_POINTS][3];
  int32_t perm_x[RT_PERLIN_POINTS];
  int32_t perm_y[RT_PERLIN_POINTS];
  int32_t perm_z[RT_PERLIN_POINTS];
} rt_perlin;

/* Image texture data: width x height texels, 3 bytes each (linear RGB, row 0 = top).  Looked up with the
 * surface coordinates the reference's primitives compute (sphere: Sphere.cpp:136-140, quad: Plane.cpp:93-104):
 * texel (int(u * width), int((1 - v) * height)), clamped; colour = byte / 255 ("Ray Tracing: The Next Week"
 * image_texture, which the reference's texture set stops short of). */
typedef struct rt_image {
  int32_t width;
  int32_t height;
  const uint8_t *rgb;
} rt_image;

/* Light-sampling proxy (geometry only; main.cpp:57-61).  Sphere: a=center, radius.  Quad:
 * a=corner, b=u, c=v. */
typedef struct rt_light {
  int32_t shape; /* RT_SHAPE_* */
  int32_t xform;
  double a[3];
  double b[3];
  double c[3];
  double radius;
} rt_light;

typedef struct rt_scene_desc {
  int32_t n_spheres;
  int32_t n_quads;
  int32_t n_xform_ops;
  int32_t n_xforms;
  int32_t n_media;
  int32_t n_materials;
  int32_t n_textures;
  int32_t n_perlins;
  int32_t n_lights;
  int32_t n_objects; /* number of top-level world objects */
  const rt_sphere *spheres;
  const rt_quad *quads;
  const rt_xform_op *xform_ops;
  const rt_xform *xforms;
  const rt_medium *media;
  const rt_material *materials;
  const rt_texture *textures;
  const rt_perlin *perlins;
  const rt_light *lights;
  /* appended in ABI version 1.1: image textures (zero-initialised descriptions have none) */
  int32_t n_images;
  int32_t pad_;
  const rt_image *images;
} rt_scene_desc;

/* Unified primitive ids reported by rt_trace_rays:
 *   [0, n_spheres)                       sphere i
 *   [n_spheres, n_spheres+n_quads)       quad i - n_spheres
 *   [n_spheres+n_quads, ... + n_media)   medium i - n_spheres - n_quads
 * Boundary primitives are never reported. */

/* ----------------------------------------------------------------------------------------------
 * Camera
 * -------------------------------------------------------------------------------------------- */

/* Mirror of CameraConfig (core/camera/CameraConfig.hpp:9-36), flags excluded. */
typedef struct rt_camera_config {
  int32_t image_width;
  int32_t samples_per_pixel;
  int32_t max_depth;
  int32_t pad_;
  double aspect_ratio;
  double vfov;
  double defocus_angle;
  double focus_dist;
  double lookfrom[3];
  double lookat[3];
  double vup[3];
  double background[3];
} rt_camera_config;

/* The derived quantities Camera::initialize computes and the reference passes to its kernels. */
typedef struct rt_camera {
  int32_t image_width;
  int32_t image_height;
  double center[3];
  double pixel00_loc[3];
  double pixel_delta_u[3];
  double pixel_delta_v[3];
  double defocus_disk_u[3];
  double defocus_disk_v[3];
  double defocus_angle;
  double background[3];
} rt_camera;

/* Camera::initialize (core/camera/Camera.cpp:31-73) in the same FP64 operation order.  Pure host
 * arithmetic, needs no device. */
int rt_camera_init(const rt_camera_config *config, rt_camera *out);

/* ----------------------------------------------------------------------------------------------
 * Context / scene
 * -------------------------------------------------------------------------------------------- */

typedef struct rt_context rt_context;
typedef struct rt_scene rt_scene;
typedef struct rt_film rt_film;

int rt_abi_version(void);
const char *rt_last_error(void);

/* Number of CUDA devices visible (0 when none; never an error). */
int rt_device_count(void);

int rt_context_create(int device, rt_context **out);
void rt_context_destroy(rt_context *ctx);
/* Block until all work queued on the context's stream is done. */
int rt_context_synchronize(rt_context *ctx);
/* The context's cudaStream_t as an integer (so callers can record CUDA events on it). */
uint64_t rt_context_stream(rt_context *ctx);

/* Upload the scene, bake instance transforms, build the binary tree (64 .. 65,536 primitives: SAH on the host;
 * otherwise on the device: Morton sort + Karras radix tree, or PLOC with RT_BVH=ploc) and collapse it into 4-wide
 * nodes on the device.  The description may be freed after the call returns. */
int rt_scene_create(rt_context *ctx, const rt_scene_desc *desc, rt_scene **out);
void rt_scene_destroy(rt_scene *scene);

enum { RT_BUILDER_NONE = 0,  /* 0 or 1 primitive */
       RT_BUILDER_SAH = 1,   /* binary SAH tree built on the host (64 .. 65,536 primitives) */
       RT_BUILDER_PLOC = 2,  /* device: Morton sort + parallel locally-ordered clustering */
       RT_BUILDER_LBVH = 3   /* device: Morton sort + Karras radix tree */ };
typedef struct rt_scene_info {
  int64_t n_prims;       /* BVH leaves: visible spheres + quads + media */
  int64_t n_nodes;       /* 4-wide nodes */
  int64_t node_bytes;
  int64_t prim_bytes;
  double build_ms;       /* device time of the tree build + collapse (+ the host SAH build where one ran) */
  double bounds_min[3];
  double bounds_max[3];
  int32_t builder;       /* RT_BUILDER_*: the tree that was kept */
  int32_t depth;         /* levels of the 4-wide tree; rt_scene_create keeps 3 * depth within the traversal stack (64
                          * entries): a device tree deeper than that is rebuilt with the host SAH builder */
} rt_scene_info;
int rt_scene_get_info(rt_scene *scene, rt_scene_info *out);

/* Animated scenes (absent from the reference, whose only motion is the per-ray shutter blend of a sphere's two
 * centres, Sphere.cpp:101-104): replaces spheres [first_sphere, first_sphere + n_spheres) of the description the
 * scene was created from - centres, radius, material, instance chain - and refits the BVH bottom-up.  The tree
 * keeps its topology: exact for any motion, but its quality degrades as objects travel far from where they were
 * built; rt_scene_create rebuilds (<= 2 ms for 10^6 primitives).  Boundary spheres of media cannot be updated
 * (RT_ERR_UNSUPPORTED).  Ordered after earlier work on the context's stream; returns when the update is done. */
int rt_scene_update_spheres(rt_scene *scene, int first_sphere, int n_spheres, const rt_sphere *spheres);
/* The same for quads [first_quad, first_quad + n_quads): corner, sides, material, instance chain. */
int rt_scene_update_quads(rt_scene *scene, int first_quad, int n_quads, const rt_quad *quads);

/* ----------------------------------------------------------------------------------------------
 * Closest-hit parity hook
 * -------------------------------------------------------------------------------------------- */

typedef struct rt_ray {
  double origin[3];
  double direction[3]; /* not normalised (Camera.cpp:199) */
  double time;
  double t_min;        /* reference uses Interval(0.001, INF) (Camera.cpp:242) */
  double t_max;
  uint32_t rng_pixel;  /* Philox key of the segment, used only by constant media */
  uint32_t rng_sample;
  uint32_t rng_bounce;
  uint32_t pad_;
} rt_ray;

typedef struct rt_hit {
  double t;           /* +inf on miss */
  int32_t prim;       /* unified primitive id, -1 on miss */
  int32_t object;     /* top-level world object index, -1 on miss */
  int32_t front_face; /* HitRecord::frontFace */
  int32_t pad_;
} rt_hit;

enum {
  RT_TRACE_EXACT_F64 = 0, /* FP64 primitive tests in the reference's operation order: ids and t bit-exact */
  RT_TRACE_FAST_F32 = 1   /* the FP32 path the renderer uses */
};

/* rays / hits are HOST pointers; n rays are copied in, traced by the extend kernel, hits copied out. */
int rt_trace_rays(rt_scene *scene, const rt_ray *rays, int64_t n, int mode, uint64_t seed, rt_hit *hits);

/* ----------------------------------------------------------------------------------------------
 * Film (accumulation buffer) and rendering
 * -------------------------------------------------------------------------------------------- */

/* A film is the un-normalised radiance sum per pixel (float4: r,g,b,unused), the analogue of the
 * reference's CudaColor accumulation buffer (DynamicCamera.cpp:458-517).  With n_ranks > 1 it only
 * holds the scanline tiles this rank owns: tile k (tile_rows consecutive scanlines) belongs to rank
 * k % n_ranks; owned tiles are stored compactly in tile order.  If external_accum is non-NULL it must
 * be a DEVICE pointer to rt_film_owned_pixels() float4s that the caller owns (e.g. a torch tensor);
 * otherwise the film allocates its own. */
int rt_film_create(rt_context *ctx, int width, int height, int rank, int n_ranks, int tile_rows,
                   void *external_accum, rt_film **out);
void rt_film_destroy(rt_film *film);
int rt_film_clear(rt_film *film);
int64_t rt_film_owned_pixels(const rt_film *film);
/* Same count without a device (for sizing the external buffer). */
int64_t rt_film_owned_pixels_for(int width, int height, int rank, int n_ranks, int tile_rows);
/* Device pointer (as integer) of the float4 accumulation buffer. */
uint64_t rt_film_device_ptr(rt_film *film);
int64_t rt_film_samples(const rt_film *film);

/* One progressive frame: one path per owned pixel through stratum (s_i, s_j) of a sqrt_spp x
 * sqrt_spp grid, ADDED to the film (cuda_dynamic_render_tile_wrapper semantics, whole frame).
 * Asynchronous on the context stream. */
int rt_render_accumulate(rt_scene *scene, const rt_camera *camera, rt_film *film, int s_i, int s_j,
                         int sqrt_spp, int max_depth, uint64_t seed);

/* n_strata consecutive strata (linear index s = s_j * sqrt_spp + s_i, starting at first_stratum) in one
 * wavefront pass, ADDED to the film: the progressive step of a camera that takes several samples per
 * displayed frame (e.g. one per GPU).  Asynchronous on the context stream. */
int rt_render_strata(rt_scene *scene, const rt_camera *camera, rt_film *film, int first_stratum, int n_strata,
                     int sqrt_spp, int max_depth, uint64_t seed);

/* All sqrt_spp^2 strata (cuda_static_render_wrapper semantics).  Clears the film first. */
int rt_render_static(rt_scene *scene, const rt_camera *camera, rt_film *film, int sqrt_spp,
                     int max_depth, uint64_t seed);

/* scale * sum as linear float RGB into a HOST buffer of owned_pixels*3 floats (compact tile order). */
int rt_film_read_rgb(rt_film *film, double scale, float *host_rgb);
/* to_byte(scale * sum) per channel (ColorUtility.hpp:18-23) into a HOST buffer of owned_pixels*3 bytes. */
int rt_film_resolve_rgb8(rt_film *film, double scale, uint8_t *host_rgb8);
/* Same into a DEVICE buffer (no copy back), asynchronous. */
int rt_film_resolve_rgb8_device(rt_film *film, double scale, void *device_rgb8);

/* After the per-rank compact films have been gathered rank-major into one DEVICE buffer
 * (n_ranks blocks, block r holding rank r's owned pixels as float4), scatter them into a full
 * row-major width x height float4 image on the device. */
int rt_film_scatter_gathered(rt_context *ctx, int width, int height, int n_ranks, int tile_rows,
                             const void *device_gathered, void *device_full_image);
/* The same for frames that every rank already resolved to RGB8 (rt_film_resolve_rgb8_device): 3 bytes per
 * pixel cross NVLink instead of 16 - what a displayed progressive frame needs. */
int rt_film_scatter_gathered_rgb8(rt_context *ctx, int width, int height, int n_ranks, int tile_rows,
                                  const void *device_gathered, void *device_full_image);

/* Single-process multi-GPU gather: copy each film's owned tiles into rank 0's full image over
 * NVLink peer copies.  films[r] must be rank r of n_ranks.  host_rgb (width*height*3 floats) gets
 * scale * sum. */
int rt_film_gather_p2p(rt_film **films, int n_ranks, double scale, float *host_rgb);
/* The same for a displayed frame: every film's tiles are tone-mapped on their own device (to_byte), the RGB8
 * tiles travel over NVLink, host_rgb8 (width*height*3 bytes) gets the frame - byte-identical to
 * rt_film_resolve_rgb8 on one GPU. */
int rt_film_gather_p2p_rgb8(rt_film **films, int n_ranks, double scale, uint8_t *host_rgb8);

/* ----------------------------------------------------------------------------------------------
 * Displayed frames of a multi-GPU render: tiles stored straight into the owner's image over NVLink
 * -------------------------------------------------------------------------------------------- */

/* A frame is a row-major width x height RGB8 image in the memory of one GPU (its owner, normally rank 0: the GPU
 * that shows or downloads it) plus one arrival flag per rank.  Every rank holds a HANDLE to the same memory:
 *   owner                      rt_frame_create
 *   another process            rt_frame_export on the owner -> 64 opaque bytes (a CUDA IPC handle; send them by any
 *                              means) -> rt_frame_open in the other process
 *   another GPU, same process  rt_frame_attach (peer access)
 * Per displayed frame, in the same order on every rank (replaces DynamicCamera::update_texture + the per-frame
 * full-buffer copy, DynamicCamera.cpp:280-306,519-554):
 *   every rank (owner included)  rt_film_present   to_byte(scale * sum) of the film's own tiles, stored into their
 *                                                  rows of the frame (remote ranks: NVLink stores), then the rank's
 *                                                  arrival flag; first waits until the owner consumed the frame's
 *                                                  previous content
 *   owner                        rt_frame_wait     the owner's stream waits (on the device) for all ranks
 *                                rt_frame_download copy to host memory on the frame's own copy stream (overlaps the
 *                                                  next render), then mark the frame consumed
 *                             or rt_frame_release  mark it consumed without a copy
 * All calls are asynchronous except rt_frame_download_wait.  Waits are bounded: a rank that never presents sets
 * the frame's error word (rt_frame_error) after ~3 s instead of hanging the device.  No cudaMalloc, no collective
 * and no staging copy happens per frame. */
typedef struct rt_frame rt_frame;
int rt_frame_create(rt_context *ctx, int width, int height, int n_ranks, rt_frame **out);
int rt_frame_export(rt_frame *frame, unsigned char handle[64]);
int rt_frame_open(rt_context *ctx, const unsigned char handle[64], int width, int height, int n_ranks, rt_frame **out);
int rt_frame_attach(rt_context *ctx, rt_frame *owner_frame, rt_frame **out);
void rt_frame_destroy(rt_frame *frame);
uint64_t rt_frame_device_ptr(rt_frame *frame); /* the RGB8 image, as seen from this handle's device */
int rt_film_present(rt_film *film, double scale, rt_frame *frame);
int rt_frame_wait(rt_frame *frame);
int rt_frame_release(rt_frame *frame);
int rt_frame_wait_release(rt_frame *frame); /* rt_frame_wait + rt_frame_release in one launch (a frame that stays on the device) */
int rt_frame_download(rt_frame *frame, uint8_t *host_rgb8); /* host_rgb8: width*height*3 bytes, pinned for overlap */
int rt_frame_download_wait(rt_frame *frame);                /* blocks until the download is in host memory */
int rt_frame_error(rt_frame *frame);                        /* 0 = fine, 1 = a wait timed out, < 0 = CUDA error */
/* Page-locked host memory for rt_frame_download (so that the copy runs asynchronously beside the next render). */
int rt_host_alloc(size_t bytes, void **out);
void rt_host_free(void *p);

/* ----------------------------------------------------------------------------------------------
 * Counters (tracing / profiling aid)
 * -------------------------------------------------------------------------------------------- */

typedef struct rt_counters {
  uint64_t paths;          /* camera samples generated */
  uint64_t segments;       /* ray segments traced (wavefront extend launches + tail kernel) */
  uint64_t kernel_launches;/* kernels launched by this library since the last reset */
  uint64_t tail_segments;  /* the part of `segments` traced by the tail kernel */
  uint64_t nodes_visited;  /* BVH4 node visits (4 box tests each) - counted only while rt_context_set_stats is on */
  uint64_t prim_tests;     /* leaf primitive tests                - counted only while rt_context_set_stats is on */
  uint64_t graph_launches; /* render passes submitted as one CUDA-graph launch (rt_context_set_graph) */
  uint64_t graph_instantiations; /* how often the executable graph had to be rebuilt instead of updated in place */
} rt_counters;
int rt_get_counters(rt_context *ctx, rt_counters *out);
int rt_reset_counters(rt_context *ctx);
/* Whole-pass graph launches (DynamicCamera.cpp:458-554 launches and synchronises per tile): while enabled, every
 * render pass - counter reset, generate, extend / shade per bounce, tail, accumulate - reaches the driver as ONE
 * cudaGraphLaunch.  The pass is captured from the context stream and the context's executable graph is updated in
 * place (camera, seed and stratum are kernel arguments), so nothing is re-instantiated from frame to frame.
 * Environment RT_GRAPH=1 turns it on at context creation.  Images are unchanged. */
int rt_context_set_graph(rt_context *ctx, int enable);
/* Traversal statistics: while enabled, the render passes of this context launch instrumented instantiations of
 * the extend / tail kernels that count node visits and primitive tests (a few per cent slower); the product
 * kernels carry no counting code. */
int rt_context_set_stats(rt_context *ctx, int enable);
/* Queue occupancy of the most recent render pass: out[b] = ray segments queued for bounce b (b < n, at most
 * max_depth + 1 entries are meaningful; the tail kernel keeps its paths in registers, so bounces it runs to
 * completion show only what it handed on).  Synchronises the context stream. */
int rt_get_queue_lengths(rt_context *ctx, uint32_t *out, int n);

/* ----------------------------------------------------------------------------------------------
 * Parity audit of the FP32 render path (test / measurement aid)
 * -------------------------------------------------------------------------------------------- */

/* While the audit is on, render passes run every bounce as a wavefront launch and, for EVERY ray segment the
 * FP32 extend kernel traces, also run the FP64 parity traversal (the reference's arithmetic, rt_trace_rays
 * RT_TRACE_EXACT_F64) on the very same ray; the tallies say how often the product path names a different
 * primitive than the reference's arithmetic would.  Slow (FP64 traversal of every segment); images are unchanged. */
typedef struct rt_audit {
  uint64_t segments;          /* segments compared */
  uint64_t prim_mismatch;     /* different primitive, or hit vs miss */
  uint64_t primary_segments;  /* bounce-0 (camera) rays compared */
  uint64_t primary_mismatch;
  uint64_t hit_miss_flips;    /* the part of prim_mismatch where one side missed */
  uint64_t t_rel_above_1e4;   /* same primitive, |t32 - t64| > 1e-4 * max(|t64|, 1e-3) */
  double max_rel_t_error;     /* over the segments with the same primitive */
  uint64_t rechecked;         /* segments whose FP32 answer was flagged uncertain and re-done in FP64 by the product */
} rt_audit;
typedef struct rt_audit_sample { /* one mismatching segment (the first RT_AUDIT_MAX_SAMPLES are kept) */
  float origin[3], time;
  float direction[3];
  int32_t bounce;
  int32_t fast_prim, exact_prim; /* unified primitive ids, -1 = miss */
  float fast_t;
  int32_t skip_prim;
  double exact_t;
} rt_audit_sample;
int rt_context_set_audit(rt_context *ctx, int enable);
int rt_get_audit(rt_context *ctx, rt_audit *out);            /* tallies since the audit was enabled / last reset */
int rt_get_audit_samples(rt_context *ctx, rt_audit_sample *out, int max_samples); /* returns the count, < 0 on error */

/* Per-stage device timing (profiling aid): while enabled, every kernel launch of the render loop is
 * bracketed by CUDA events on the context stream.  rt_get_stage_times synchronises, adds the elapsed
 * milliseconds and launch counts per stage since the last call and resets them. */
enum { RT_STAGE_GENERATE = 0, RT_STAGE_EXTEND = 1, RT_STAGE_SHADE = 2, RT_STAGE_ACCUMULATE = 3, RT_STAGE_TAIL = 4,
       RT_STAGE_COUNT = 5 };
int rt_context_set_stage_timing(rt_context *ctx, int enable);
int rt_get_stage_times(rt_context *ctx, double ms[RT_STAGE_COUNT], uint64_t launches[RT_STAGE_COUNT]);

#ifdef __cplusplus
}
#endif
#endif /* RT_B200_H */
