"""The oracle (oracle/rt_oracle.c) against the golden fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import oracle_lib as ol
from conftest import array_from, load_golden, struct_from
from rt_b200 import abi

SCENES = ["spheres", "spheres_textured", "cornell", "cornell_smoke", "final"]


def test_mt19937_streams(oracle):
    g = load_golden("rng")
    st = ol.ora_mt19937()
    oracle.ora_mt_seed(st, 1234)
    got = np.array([oracle.ora_mt_canonical(st) for _ in range(256)])
    assert np.array_equal(got, g["canonical"])  # random_double(), Utility.hpp:22-25
    for key in g.files:
        if not key.startswith("int_"):
            continue
        lo, hi = map(int, key.split("_")[1:])
        oracle.ora_mt_seed(st, 1234 + hi)
        got = np.array([oracle.ora_mt_uniform_int(st, lo, hi) for _ in range(128)], dtype=np.int64)
        assert np.array_equal(got, g[key]), key  # random_int(), Utility.hpp:34-37


def test_mt19937_known_answer(oracle):
    # the C++ standard requires the 10000th output of a default-seeded mt19937 to be 4123659995
    st = ol.ora_mt19937()
    oracle.ora_mt_seed(st, 5489)
    v = 0
    for _ in range(10000):
        v = oracle.ora_mt_next(st)
    assert v == 4123659995


@pytest.mark.parametrize("ctr,key,want", [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
])
def test_philox_known_answers(oracle, ctr, key, want):
    # Random123's published known-answer vectors for philox4x32-10
    out = (C.c_uint32 * 4)()
    oracle.ora_philox4x32_10((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
    assert tuple(out) == want


def test_to_byte(oracle):
    g = load_golden("tonemap")
    got = np.array([oracle.ora_to_byte(float(x)) for x in g["x"]], dtype=np.uint8)
    assert np.array_equal(got, g["byte"])  # ColorUtility.hpp:18-26


def test_sphere_uv_table(oracle):
    # the reference's only in-tree known-answer data: the comment table in Sphere.cpp:129-134
    table = {(1, 0, 0): (0.50, 0.50), (-1, 0, 0): (0.00, 0.50), (0, 1, 0): (0.50, 1.00), (0, -1, 0): (0.50, 0.00),
             (0, 0, 1): (0.25, 0.50), (0, 0, -1): (0.75, 0.50)}
    for n, (u, v) in table.items():
        gu, gv = C.c_double(), C.c_double()
        oracle.ora_sphere_uv((C.c_double * 3)(*map(float, n)), C.byref(gu), C.byref(gv))
        assert abs(gv.value - v) < 1e-12
        if n[1] == 0:  # u is arbitrary at the poles
            assert abs(gu.value % 1.0 - u % 1.0) < 1e-12


def test_builtin_scenes_match_reference(scene_index, host_scenes):
    """The product's scene generators emit the reference's scenes bit for bit (checksums of the arrays the
    reference harness recorded while constructing the same objects from reference classes)."""
    for key, meta in scene_index.items():
        hs = host_scenes(meta["builtin"], meta["p0"], meta["p1"], meta["seed"])
        desc = hs.desc.contents
        for f, want in meta["counts"].items():
            assert getattr(desc, f) == want, (key, f)
        got = [hashlib.sha256(p).hexdigest() for p in ol.desc_bytes(desc)]
        assert got == meta["sha256"], key


@pytest.mark.parametrize("name", SCENES)
def test_camera_and_primary_rays(oracle, scene_index, name):
    g = load_golden(name)
    cfg = struct_from(abi.rt_camera_config, g["camera_config"])
    cam = abi.rt_camera()
    oracle.ora_camera_init(C.byref(cfg), C.byref(cam))
    assert bytes(cam) == g["camera"].tobytes()  # Camera::initialize, Camera.cpp:31-73
    # the product's own camera set-up (host arithmetic inside librt_b200.so) must agree as well
    lib = abi.load_library()
    cam2 = abi.rt_camera()
    assert lib.rt_camera_init(C.byref(cfg), C.byref(cam2)) == 0
    assert bytes(cam2) == g["camera"].tobytes()
    n = cam.image_width * cam.image_height
    rays = (abi.rt_ray * n)()
    oracle.ora_primary_rays(C.byref(cfg), ol.ORA_RNG_MT19937, ol.ORA_SAMPLER_REJECTION, 77, 1, 0, rays)
    want = ol.rays_to_numpy(array_from(abi.rt_ray, g["primary_rays"]))
    got = ol.rays_to_numpy(rays)
    for f in ("origin", "direction", "time"):
        assert np.array_equal(got[f], want[f]), f  # Camera::get_ray, Camera.cpp:186-205


@pytest.mark.parametrize("name", SCENES)
def test_closest_hits(oracle, scene_index, host_scenes, name):
    g = load_golden(name)
    meta = scene_index[name]
    hs = host_scenes(meta["builtin"], meta["p0"], meta["p1"], meta["seed"])
    osc = oracle.ora_scene_create(hs.desc)
    rays = array_from(abi.rt_ray, g["primary_rays"])
    has_media = meta["counts"]["n_media"] > 0
    for use_bvh in (0, 1):
        if has_media and use_bvh:
            continue  # media draw inside hit(): the reference's SAH tree visits them in another order
        want = ol.hits_to_numpy(array_from(abi.rt_hit, g[f"primary_hits_bvh{use_bvh}"]))
        got = (abi.rt_hit * len(rays))()
        oracle.ora_trace(osc, rays, len(rays), use_bvh, ol.ORA_RNG_MT19937, 5, got)
        got = ol.hits_to_numpy(got)
        assert np.array_equal(got["t"], want["t"])
        assert np.array_equal(got["object"], want["object"])
        assert np.array_equal(got["front_face"], want["front_face"])
    if "segment_rays" in g.files:  # every segment of a small render, answered by the reference
        rays = array_from(abi.rt_ray, g["segment_rays"])
        want = ol.hits_to_numpy(array_from(abi.rt_hit, g["segment_hits"]))
        for use_bvh in (0, 1):
            got = (abi.rt_hit * len(rays))()
            oracle.ora_trace(osc, rays, len(rays), use_bvh, ol.ORA_RNG_MT19937, 5, got)
            got = ol.hits_to_numpy(got)
            assert np.array_equal(got["t"], want["t"])
            assert np.array_equal(got["object"], want["object"])
            assert np.array_equal(got["front_face"], want["front_face"])
    oracle.ora_scene_destroy(osc)


@pytest.mark.parametrize("name", SCENES)
def test_renders_bit_exact(oracle, scene_index, host_scenes, name):
    """Camera::ray_color over a whole (small) image: materials, textures, light sampling, media."""
    g = load_golden(name)
    meta = scene_index[name]
    hs = host_scenes(meta["builtin"], meta["p0"], meta["p1"], meta["seed"])
    osc = oracle.ora_scene_create(hs.desc)
    cfg = struct_from(abi.rt_camera_config, g["camera_config"])
    cam = abi.rt_camera()
    oracle.ora_camera_init(C.byref(cfg), C.byref(cam))
    n = cam.image_width * cam.image_height
    has_media = meta["counts"]["n_media"] > 0
    for use_bvh in (0, 1):
        if has_media and use_bvh:
            continue
        img = (C.c_double * (n * 3))()
        cnt = ol.ora_counters()
        oracle.ora_render(osc, C.byref(cfg), ol.ORA_RNG_MT19937, ol.ORA_SAMPLER_REJECTION, 99, use_bvh, 0,
                          cam.image_height, -1, img, C.byref(cnt))
        assert np.array_equal(np.frombuffer(img, dtype=np.float64), g[f"render_bvh{use_bvh}"], equal_nan=True)
        assert cnt.segments == int(g[f"render_segments_bvh{use_bvh}"][0])
    oracle.ora_scene_destroy(osc)
