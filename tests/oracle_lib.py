"""ctypes bindings of the TEST-ONLY checkers: oracle/liboracle.so (C restatement) and, when present,
oracle/_ref/libref_harness.so (the unmodified reference compiled by oracle/Makefile)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))

from rt_b200 import abi  # noqa: E402

ORACLE_PATH = os.path.join(REPO, "oracle", "liboracle.so")
REF_PATH = os.path.join(REPO, "oracle", "_ref", "libref_harness.so")
REF_GPU_PATH = os.path.join(REPO, "oracle", "_ref_gpu", "libref_gpu.so")  # the reference's own CUDA path (comparator)

ORA_RNG_MT19937, ORA_RNG_PHILOX = 0, 1
ORA_SAMPLER_REJECTION, ORA_SAMPLER_POLAR = 0, 1

P = C.POINTER


class ora_counters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("node_tests", C.c_uint64),
                ("sphere_tests", C.c_uint64), ("quad_tests", C.c_uint64), ("rng_draws", C.c_uint64)]


class ora_mt19937(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int)]


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_PATH):
            subprocess.check_call(["make", "-C", os.path.join(REPO, "oracle"), "liboracle.so"])
        lib = C.CDLL(ORACLE_PATH)
        lib.ora_scene_create.restype = C.c_void_p
        lib.ora_scene_create.argtypes = [P(abi.rt_scene_desc)]
        lib.ora_scene_destroy.argtypes = [C.c_void_p]
        lib.ora_camera_init.argtypes = [P(abi.rt_camera_config), P(abi.rt_camera)]
        lib.ora_primary_rays.argtypes = [P(abi.rt_camera_config), C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                         P(abi.rt_ray)]
        lib.ora_trace.argtypes = [C.c_void_p, P(abi.rt_ray), C.c_int64, C.c_int, C.c_int, C.c_uint64, P(abi.rt_hit)]
        lib.ora_render.restype = C.c_double
        lib.ora_render.argtypes = [C.c_void_p, P(abi.rt_camera_config), C.c_int, C.c_int, C.c_uint64, C.c_int,
                                   C.c_int, C.c_int, C.c_int, P(C.c_double), P(ora_counters)]
        lib.ora_to_byte.restype = C.c_int
        lib.ora_to_byte.argtypes = [C.c_double]
        lib.ora_sphere_uv.argtypes = [P(C.c_double), P(C.c_double), P(C.c_double)]
        lib.ora_mt_seed.argtypes = [P(ora_mt19937), C.c_uint32]
        lib.ora_mt_next.restype = C.c_uint32
        lib.ora_mt_next.argtypes = [P(ora_mt19937)]
        lib.ora_mt_canonical.restype = C.c_double
        lib.ora_mt_canonical.argtypes = [P(ora_mt19937)]
        lib.ora_mt_uniform_int.restype = C.c_int
        lib.ora_mt_uniform_int.argtypes = [P(ora_mt19937), C.c_int, C.c_int]
        lib.ora_philox4x32_10.argtypes = [P(C.c_uint32), P(C.c_uint32), P(C.c_uint32)]
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_PATH)
        lib.ref_scene_build.restype = C.c_void_p
        lib.ref_scene_build.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int]
        lib.ref_scene_free.argtypes = [C.c_void_p]
        lib.ref_scene_desc.restype = P(abi.rt_scene_desc)
        lib.ref_scene_desc.argtypes = [C.c_void_p]
        lib.ref_scene_camera_config.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, P(abi.rt_camera_config)]
        lib.ref_scene_set_aspect.argtypes = [C.c_void_p, C.c_double]
        lib.ref_camera_init.argtypes = [P(abi.rt_camera_config), P(abi.rt_camera)]
        lib.ref_primary_rays.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, P(abi.rt_ray)]
        lib.ref_trace.argtypes = [C.c_void_p, P(abi.rt_ray), C.c_int64, C.c_int, P(abi.rt_hit)]
        lib.ref_seed.argtypes = [C.c_uint64]
        lib.ref_random_double.restype = C.c_double
        lib.ref_random_int.restype = C.c_int
        lib.ref_random_int.argtypes = [C.c_int, C.c_int]
        lib.ref_render.restype = C.c_double
        lib.ref_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, P(C.c_double), P(C.c_uint64)]
        lib.ref_to_byte.restype = C.c_int
        lib.ref_to_byte.argtypes = [C.c_double]
        lib.ref_hardware_threads.restype = C.c_int
        _ref = lib
    return _ref


def have_ref_gpu():
    return os.path.exists(REF_GPU_PATH)


def ref_gpu():
    """The reference compiled with -DUSE_CUDA (oracle/Makefile `gpu`).  Load it in a process of its own: it uses the
    default stream, process-global scene state and device recursion, and a fault in it poisons the CUDA context."""
    lib = C.CDLL(REF_GPU_PATH)
    lib.ref_gpu_frame.restype = C.c_int
    lib.ref_gpu_frame.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  P(C.c_double), P(C.c_double), C.c_char_p, C.c_int]
    return lib


# ---------------------------------------------------------------------------------------------------
# numpy views of the ABI structs
# ---------------------------------------------------------------------------------------------------
def struct_array(ctype, n):
    return (ctype * n)()


def as_numpy(arr):
    """Structured numpy view of a ctypes array of Structures."""
    return np.ctypeslib.as_array(arr)


def rays_to_numpy(rays):
    a = np.frombuffer(rays, dtype=np.dtype([("origin", "<f8", 3), ("direction", "<f8", 3), ("time", "<f8"),
                                            ("t_min", "<f8"), ("t_max", "<f8"), ("rng_pixel", "<u4"),
                                            ("rng_sample", "<u4"), ("rng_bounce", "<u4"), ("pad_", "<u4")]))
    return a


def hits_to_numpy(hits):
    return np.frombuffer(hits, dtype=np.dtype([("t", "<f8"), ("prim", "<i4"), ("object", "<i4"),
                                               ("front_face", "<i4"), ("pad_", "<i4")]))


def desc_bytes(desc):
    """Concatenated raw bytes of every array of a scene description (for checksums / equality)."""
    d = desc
    parts = []
    for name, cnt in (("spheres", d.n_spheres), ("quads", d.n_quads), ("xform_ops", d.n_xform_ops),
                      ("xforms", d.n_xforms), ("media", d.n_media), ("materials", d.n_materials),
                      ("textures", d.n_textures), ("perlins", d.n_perlins), ("lights", d.n_lights)):
        ptr = getattr(d, name)
        if cnt:
            parts.append(C.string_at(ptr, cnt * C.sizeof(ptr._type_)))
        else:
            parts.append(b"")
    return parts
