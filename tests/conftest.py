import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, "real-time-ray-tracing-engine_b200"))

import oracle_lib as ol  # noqa: E402
from rt_b200 import abi  # noqa: E402

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices():
    """Devices the product library sees.  A missing / unloadable library is NOT a reason to skip: the gpu tests
    then run and fail loudly (on the GPU box that is a broken build, not a missing GPU)."""
    try:
        return int(abi.load_library().rt_device_count())
    except Exception:
        return 1


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a GPU skips the gpu-marked tests instead of failing them (the
    product has no CPU fallback, so they cannot run there)."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the CUDA path is the product, there is no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    o = ol.oracle()
    o.ora_ray_log.argtypes = [C.POINTER(abi.rt_ray), C.c_int64]
    o.ora_ray_log_count.restype = C.c_int64
    return o


@pytest.fixture(scope="session")
def scene_index():
    with open(os.path.join(GOLDEN, "scenes.json")) as f:
        return json.load(f)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def struct_from(ctype, arr):
    return ctype.from_buffer_copy(arr.tobytes())


def array_from(ctype, arr):
    n = arr.size // C.sizeof(ctype)
    return (ctype * n).from_buffer_copy(arr.tobytes())


@pytest.fixture(scope="session")
def host_scenes():
    """Built-in scenes from the product's host library, cached per (name, p0, p1)."""
    from rt_b200 import host

    cache = {}

    def get(name, p0=0, p1=-1, seed=1234):
        key = (name, p0, p1, seed)
        if key not in cache:
            cache[key] = host.HostScene.builtin(name, seed, p0, p1)
        return cache[key]

    return get


def oracle_segments(o, osc, cfg, rng_kind, sampler, seed, use_bvh, single_stratum=-1):
    """Render with the oracle and return (image[n,3], rays ctypes array of every traced segment)."""
    cam = abi.rt_camera()
    o.ora_camera_init(C.byref(cfg), C.byref(cam))
    n = cam.image_width * cam.image_height
    spp = 1 if single_stratum >= 0 else int(np.sqrt(cfg.samples_per_pixel)) ** 2
    cap = n * spp * cfg.max_depth
    log = (abi.rt_ray * cap)()
    o.ora_ray_log(log, cap)
    img = (C.c_double * (n * 3))()
    cnt = ol.ora_counters()
    o.ora_render(osc, C.byref(cfg), rng_kind, sampler, seed, use_bvh, 0, cam.image_height, single_stratum, img,
                 C.byref(cnt))
    n_seg = o.ora_ray_log_count()
    o.ora_ray_log(None, 0)
    rays = (abi.rt_ray * n_seg).from_buffer_copy(bytes(log)[: n_seg * C.sizeof(abi.rt_ray)])
    return np.frombuffer(img, dtype=np.float64).reshape(-1, 3).copy(), rays, cnt
