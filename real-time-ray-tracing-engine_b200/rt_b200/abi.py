"""ctypes mirror of include/rt_b200.h (the C ABI of librt_b200.so).

Only plain structs and the library loader live here.  The loader fails loudly when the CUDA library
has not been built: there is no CPU fallback (include/rt_b200.h, RT_ERR_NO_DEVICE).
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO_DIR = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(PKG_DIR, "csrc", "librt_b200.so")  # override: A/B builds
HOST_LIB_PATH = os.path.join(PKG_DIR, "host", "librt_host.so")

RT_B200_ABI_VERSION = 2
RT_OK, RT_ERR_INVALID, RT_ERR_CUDA, RT_ERR_NO_DEVICE, RT_ERR_UNSUPPORTED = range(5)
RT_MAT_LAMBERTIAN, RT_MAT_METAL, RT_MAT_DIELECTRIC, RT_MAT_DIFFUSE_LIGHT, RT_MAT_ISOTROPIC = range(5)
RT_TEX_SOLID, RT_TEX_CHECKER, RT_TEX_NOISE, RT_TEX_IMAGE = range(4)
RT_XF_TRANSLATE, RT_XF_ROTATE_Y = range(2)
RT_SHAPE_SPHERE, RT_SHAPE_QUAD = range(2)
RT_PRIM_BOUNDARY = 1
RT_PERLIN_POINTS = 256
RT_TRACE_EXACT_F64, RT_TRACE_FAST_F32 = range(2)
RT_BUILDER_NONE, RT_BUILDER_SAH, RT_BUILDER_PLOC, RT_BUILDER_LBVH = range(4)

d3 = C.c_double * 3


class rt_sphere(C.Structure):
    _fields_ = [("center0", d3), ("center_dir", d3), ("radius", C.c_double), ("material", C.c_int32),
                ("xform", C.c_int32), ("object", C.c_int32), ("flags", C.c_int32)]


class rt_quad(C.Structure):
    _fields_ = [("corner", d3), ("u", d3), ("v", d3), ("material", C.c_int32), ("xform", C.c_int32),
                ("object", C.c_int32), ("flags", C.c_int32)]


class rt_xform_op(C.Structure):
    _fields_ = [("type", C.c_int32), ("pad_", C.c_int32), ("offset", d3), ("angle_deg", C.c_double),
                ("sin_theta", C.c_double), ("cos_theta", C.c_double)]


class rt_xform(C.Structure):
    _fields_ = [("first_op", C.c_int32), ("n_ops", C.c_int32)]


class rt_medium(C.Structure):
    _fields_ = [("density", C.c_double), ("shape", C.c_int32), ("first_prim", C.c_int32), ("n_prims", C.c_int32),
                ("material", C.c_int32), ("object", C.c_int32), ("pad_", C.c_int32)]


class rt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("texture", C.c_int32), ("albedo", d3), ("fuzz", C.c_double),
                ("ior", C.c_double)]


class rt_texture(C.Structure):
    _fields_ = [("type", C.c_int32), ("even", C.c_int32), ("odd", C.c_int32), ("perlin", C.c_int32),
                ("color", d3), ("scale", C.c_double)]


class rt_perlin(C.Structure):
    _fields_ = [("rand_vec", (C.c_double * 3) * RT_PERLIN_POINTS), ("perm_x", C.c_int32 * RT_PERLIN_POINTS),
                ("perm_y", C.c_int32 * RT_PERLIN_POINTS), ("perm_z", C.c_int32 * RT_PERLIN_POINTS)]


class rt_light(C.Structure):
    _fields_ = [("shape", C.c_int32), ("xform", C.c_int32), ("a", d3), ("b", d3), ("c", d3),
                ("radius", C.c_double)]


class rt_image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("rgb", C.POINTER(C.c_uint8))]


class rt_scene_desc(C.Structure):
    _fields_ = [("n_spheres", C.c_int32), ("n_quads", C.c_int32), ("n_xform_ops", C.c_int32),
                ("n_xforms", C.c_int32), ("n_media", C.c_int32), ("n_materials", C.c_int32),
                ("n_textures", C.c_int32), ("n_perlins", C.c_int32), ("n_lights", C.c_int32),
                ("n_objects", C.c_int32),
                ("spheres", C.POINTER(rt_sphere)), ("quads", C.POINTER(rt_quad)),
                ("xform_ops", C.POINTER(rt_xform_op)), ("xforms", C.POINTER(rt_xform)),
                ("media", C.POINTER(rt_medium)), ("materials", C.POINTER(rt_material)),
                ("textures", C.POINTER(rt_texture)), ("perlins", C.POINTER(rt_perlin)),
                ("lights", C.POINTER(rt_light)),
                ("n_images", C.c_int32), ("pad_", C.c_int32), ("images", C.POINTER(rt_image))]


class rt_camera_config(C.Structure):
    _fields_ = [("image_width", C.c_int32), ("samples_per_pixel", C.c_int32), ("max_depth", C.c_int32),
                ("pad_", C.c_int32), ("aspect_ratio", C.c_double), ("vfov", C.c_double),
                ("defocus_angle", C.c_double), ("focus_dist", C.c_double), ("lookfrom", d3), ("lookat", d3),
                ("vup", d3), ("background", d3)]


class rt_camera(C.Structure):
    _fields_ = [("image_width", C.c_int32), ("image_height", C.c_int32), ("center", d3), ("pixel00_loc", d3),
                ("pixel_delta_u", d3), ("pixel_delta_v", d3), ("defocus_disk_u", d3), ("defocus_disk_v", d3),
                ("defocus_angle", C.c_double), ("background", d3)]


class rt_scene_info(C.Structure):
    _fields_ = [("n_prims", C.c_int64), ("n_nodes", C.c_int64), ("node_bytes", C.c_int64),
                ("prim_bytes", C.c_int64), ("build_ms", C.c_double), ("bounds_min", d3), ("bounds_max", d3),
                ("builder", C.c_int32), ("depth", C.c_int32)]


class rt_ray(C.Structure):
    _fields_ = [("origin", d3), ("direction", d3), ("time", C.c_double), ("t_min", C.c_double),
                ("t_max", C.c_double), ("rng_pixel", C.c_uint32), ("rng_sample", C.c_uint32),
                ("rng_bounce", C.c_uint32), ("pad_", C.c_uint32)]


class rt_hit(C.Structure):
    _fields_ = [("t", C.c_double), ("prim", C.c_int32), ("object", C.c_int32), ("front_face", C.c_int32),
                ("pad_", C.c_int32)]


class rt_counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("tail_segments", C.c_uint64), ("nodes_visited", C.c_uint64), ("prim_tests", C.c_uint64),
                ("graph_launches", C.c_uint64), ("graph_instantiations", C.c_uint64)]


class rt_audit(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("prim_mismatch", C.c_uint64), ("primary_segments", C.c_uint64),
                ("primary_mismatch", C.c_uint64), ("hit_miss_flips", C.c_uint64), ("t_rel_above_1e4", C.c_uint64),
                ("max_rel_t_error", C.c_double), ("rechecked", C.c_uint64)]


class rt_audit_sample(C.Structure):
    _fields_ = [("origin", C.c_float * 3), ("time", C.c_float), ("direction", C.c_float * 3), ("bounce", C.c_int32),
                ("fast_prim", C.c_int32), ("exact_prim", C.c_int32), ("fast_t", C.c_float), ("skip_prim", C.c_int32),
                ("exact_t", C.c_double)]


P = C.POINTER
VOIDPP = P(C.c_void_p)

# name -> (restype, argtypes): every symbol include/rt_b200.h declares
RT_B200_SYMBOLS = {
    "rt_camera_init": (C.c_int, [P(rt_camera_config), P(rt_camera)]),
    "rt_abi_version": (C.c_int, []),
    "rt_last_error": (C.c_char_p, []),
    "rt_device_count": (C.c_int, []),
    "rt_context_create": (C.c_int, [C.c_int, VOIDPP]),
    "rt_context_destroy": (None, [C.c_void_p]),
    "rt_context_synchronize": (C.c_int, [C.c_void_p]),
    "rt_context_stream": (C.c_uint64, [C.c_void_p]),
    "rt_scene_create": (C.c_int, [C.c_void_p, P(rt_scene_desc), VOIDPP]),
    "rt_scene_destroy": (None, [C.c_void_p]),
    "rt_scene_get_info": (C.c_int, [C.c_void_p, P(rt_scene_info)]),
    "rt_scene_update_spheres": (C.c_int, [C.c_void_p, C.c_int, C.c_int, P(rt_sphere)]),
    "rt_scene_update_quads": (C.c_int, [C.c_void_p, C.c_int, C.c_int, P(rt_quad)]),
    "rt_trace_rays": (C.c_int, [C.c_void_p, P(rt_ray), C.c_int64, C.c_int, C.c_uint64, P(rt_hit)]),
    "rt_film_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, VOIDPP]),
    "rt_film_destroy": (None, [C.c_void_p]),
    "rt_film_clear": (C.c_int, [C.c_void_p]),
    "rt_film_owned_pixels": (C.c_int64, [C.c_void_p]),
    "rt_film_owned_pixels_for": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rt_film_device_ptr": (C.c_uint64, [C.c_void_p]),
    "rt_film_samples": (C.c_int64, [C.c_void_p]),
    "rt_render_accumulate": (C.c_int, [C.c_void_p, P(rt_camera), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_uint64]),
    "rt_render_strata": (C.c_int, [C.c_void_p, P(rt_camera), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_uint64]),
    "rt_render_static": (C.c_int, [C.c_void_p, P(rt_camera), C.c_void_p, C.c_int, C.c_int, C.c_uint64]),
    "rt_film_read_rgb": (C.c_int, [C.c_void_p, C.c_double, P(C.c_float)]),
    "rt_film_resolve_rgb8": (C.c_int, [C.c_void_p, C.c_double, P(C.c_uint8)]),
    "rt_film_resolve_rgb8_device": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "rt_film_scatter_gathered": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "rt_film_scatter_gathered_rgb8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                                C.c_void_p]),
    "rt_film_gather_p2p": (C.c_int, [P(C.c_void_p), C.c_int, C.c_double, P(C.c_float)]),
    "rt_film_gather_p2p_rgb8": (C.c_int, [P(C.c_void_p), C.c_int, C.c_double, P(C.c_uint8)]),
    "rt_frame_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, VOIDPP]),
    "rt_frame_export": (C.c_int, [C.c_void_p, P(C.c_ubyte)]),
    "rt_frame_open": (C.c_int, [C.c_void_p, P(C.c_ubyte), C.c_int, C.c_int, C.c_int, VOIDPP]),
    "rt_frame_attach": (C.c_int, [C.c_void_p, C.c_void_p, VOIDPP]),
    "rt_frame_destroy": (None, [C.c_void_p]),
    "rt_frame_device_ptr": (C.c_uint64, [C.c_void_p]),
    "rt_film_present": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "rt_frame_wait": (C.c_int, [C.c_void_p]),
    "rt_frame_release": (C.c_int, [C.c_void_p]),
    "rt_frame_wait_release": (C.c_int, [C.c_void_p]),
    "rt_frame_download": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rt_frame_download_wait": (C.c_int, [C.c_void_p]),
    "rt_frame_error": (C.c_int, [C.c_void_p]),
    "rt_host_alloc": (C.c_int, [C.c_size_t, VOIDPP]),
    "rt_host_free": (None, [C.c_void_p]),
    "rt_get_counters": (C.c_int, [C.c_void_p, P(rt_counters)]),
    "rt_reset_counters": (C.c_int, [C.c_void_p]),
    "rt_context_set_graph": (C.c_int, [C.c_void_p, C.c_int]),
    "rt_context_set_stats": (C.c_int, [C.c_void_p, C.c_int]),
    "rt_get_queue_lengths": (C.c_int, [C.c_void_p, P(C.c_uint32), C.c_int]),
    "rt_context_set_audit": (C.c_int, [C.c_void_p, C.c_int]),
    "rt_get_audit": (C.c_int, [C.c_void_p, P(rt_audit)]),
    "rt_get_audit_samples": (C.c_int, [C.c_void_p, P(rt_audit_sample), C.c_int]),
    "rt_context_set_stage_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "rt_get_stage_times": (C.c_int, [C.c_void_p, P(C.c_double), P(C.c_uint64)]),
}


class RtError(RuntimeError):
    pass


_lib = None


def load_library(path=None):
    """Load librt_b200.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RtError(f"{p} not found: build it with `make -C {os.path.dirname(p)}` "
                      "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in RT_B200_SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
        fn.restype = res
        fn.argtypes = args
    if lib.rt_abi_version() != RT_B200_ABI_VERSION:
        raise RtError("librt_b200.so ABI version mismatch")
    if path is None:
        _lib = lib
    return lib


def check(lib, status, what):
    if status != RT_OK:
        msg = lib.rt_last_error()
        raise RtError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")
