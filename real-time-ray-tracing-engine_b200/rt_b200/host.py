"""ctypes binding of librt_host.so (include/rt_host.h): scene construction, JSON scenes, CLI, PPM."""
import ctypes as C
import os

from . import abi

_lib = None


def load_host_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(abi.HOST_LIB_PATH):
        raise abi.RtError(f"{abi.HOST_LIB_PATH} not found: build it with `make -C {os.path.dirname(abi.HOST_LIB_PATH)}`")
    lib = C.CDLL(abi.HOST_LIB_PATH)
    lib.rth_last_error.restype = C.c_char_p
    lib.rth_scene_builtin.restype = C.c_void_p
    lib.rth_scene_builtin.argtypes = [C.c_char_p, C.c_uint64, C.c_int, C.c_int]
    lib.rth_scene_free.argtypes = [C.c_void_p]
    lib.rth_scene_desc.restype = C.POINTER(abi.rt_scene_desc)
    lib.rth_scene_desc.argtypes = [C.c_void_p]
    lib.rth_scene_camera.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(abi.rt_camera_config)]
    lib.rth_write_ppm_p3.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint8)]
    if hasattr(lib, "rth_scene_load_json"):
        lib.rth_scene_load_json.restype = C.c_void_p
        lib.rth_scene_load_json.argtypes = [C.c_char_p]
        lib.rth_scene_save_json.argtypes = [C.c_void_p, C.c_char_p]
    _lib = lib
    return lib


class HostScene:
    """A flat scene description held by the host library (built-in generator or JSON file)."""

    def __init__(self, handle):
        self._lib = load_host_library()
        if not handle:
            raise abi.RtError("scene: " + self._lib.rth_last_error().decode())
        self._h = C.c_void_p(handle)

    @classmethod
    def builtin(cls, name, seed=1234, p0=0, p1=-1):
        lib = load_host_library()
        return cls(lib.rth_scene_builtin(name.encode(), seed, p0, p1))

    @classmethod
    def from_json(cls, path):
        lib = load_host_library()
        return cls(lib.rth_scene_load_json(os.fsencode(path)))

    def save_json(self, path):
        if self._lib.rth_scene_save_json(self._h, os.fsencode(path)) != 0:
            raise abi.RtError("save_json: " + self._lib.rth_last_error().decode())

    @property
    def desc(self):
        return self._lib.rth_scene_desc(self._h)

    def camera_config(self, image_width, samples_per_pixel, max_depth):
        cfg = abi.rt_camera_config()
        self._lib.rth_scene_camera(self._h, image_width, samples_per_pixel, max_depth, C.byref(cfg))
        return cfg

    def close(self):
        if self._h:
            self._lib.rth_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_ppm_p3(path, width, height, rgb8):
    """rgb8: bytes-like / numpy uint8 array of width*height*3."""
    import numpy as np

    arr = np.ascontiguousarray(rgb8, dtype=np.uint8).reshape(-1)
    assert arr.size == width * height * 3
    lib = load_host_library()
    if lib.rth_write_ppm_p3(os.fsencode(path), width, height, arr.ctypes.data_as(C.POINTER(C.c_uint8))) != 0:
        raise abi.RtError("write_ppm_p3: " + lib.rth_last_error().decode())
