"""The oracle against the live reference (oracle/_ref/libref_harness.so = unmodified reference sources
compiled by oracle/Makefile), at other sizes and seeds than the committed fixtures.  CPU only; skipped
where the harness has not been built (it cannot be built where /root/reference does not exist, but the
prebuilt library travels with the repository snapshot)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from rt_b200 import abi

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libref_harness.so not built")

CASES = [("spheres", 11, 0, 96), ("spheres_textured", 8, 0, 64), ("cornell", 0, 0, 56), ("cornell_smoke", 0, 0, 56),
         ("final", 5, 64, 72)]


@pytest.mark.parametrize("name,p0,p1,width", CASES)
def test_oracle_reproduces_reference(oracle, host_scenes, name, p0, p1, width):
    r = ol.ref()
    h = r.ref_scene_build(name.encode(), 4321, p0, p1)
    desc = r.ref_scene_desc(h).contents
    # scene generators: product host library == reference classes
    mine = host_scenes(name, p0, p1, 4321).desc.contents
    assert ol.desc_bytes(mine) == ol.desc_bytes(desc)
    cfg = abi.rt_camera_config()
    r.ref_scene_camera_config(h, width, 9, 6, cfg)
    cam_r, cam_o = abi.rt_camera(), abi.rt_camera()
    r.ref_camera_init(cfg, cam_r)
    oracle.ora_camera_init(C.byref(cfg), C.byref(cam_o))
    assert bytes(cam_r) == bytes(cam_o)
    n = cam_r.image_width * cam_r.image_height
    rays_r, rays_o = (abi.rt_ray * n)(), (abi.rt_ray * n)()
    r.ref_primary_rays(h, width, 9, 31, 2, 1, rays_r)
    oracle.ora_primary_rays(C.byref(cfg), ol.ORA_RNG_MT19937, ol.ORA_SAMPLER_REJECTION, 31, 2, 1, rays_o)
    a, b = ol.rays_to_numpy(rays_r), ol.rays_to_numpy(rays_o)
    for f in ("origin", "direction", "time"):
        assert np.array_equal(a[f], b[f])
    osc = oracle.ora_scene_create(C.byref(desc))
    has_media = desc.n_media > 0
    for use_bvh in (0, 1):
        if has_media and use_bvh:
            continue
        hr, ho = (abi.rt_hit * n)(), (abi.rt_hit * n)()
        r.ref_seed(8)
        r.ref_trace(h, rays_r, n, use_bvh, hr)
        oracle.ora_trace(osc, rays_r, n, use_bvh, ol.ORA_RNG_MT19937, 8, ho)
        a, b = ol.hits_to_numpy(hr), ol.hits_to_numpy(ho)
        assert np.array_equal(a["t"], b["t"]) and np.array_equal(a["object"], b["object"])
        assert np.array_equal(a["front_face"], b["front_face"])
        img_r, img_o = (C.c_double * (n * 3))(), (C.c_double * (n * 3))()
        seg = C.c_uint64()
        cnt = ol.ora_counters()
        r.ref_render(h, width, 9, 6, 17, use_bvh, 1, 0, cam_r.image_height, -1, img_r, C.byref(seg))
        oracle.ora_render(osc, C.byref(cfg), ol.ORA_RNG_MT19937, ol.ORA_SAMPLER_REJECTION, 17, use_bvh, 0,
                          cam_r.image_height, -1, img_o, C.byref(cnt))
        assert np.array_equal(np.frombuffer(img_r, dtype=np.float64), np.frombuffer(img_o, dtype=np.float64),
                              equal_nan=True)
        assert seg.value == cnt.segments
    # one dynamic-mode frame (single stratum, un-normalised; DynamicCamera.cpp:103-171)
    img_r, img_o = (C.c_double * (n * 3))(), (C.c_double * (n * 3))()
    r.ref_render(h, width, 9, 6, 3, 0, 1, 0, cam_r.image_height, 4, img_r, None)
    oracle.ora_render(osc, C.byref(cfg), ol.ORA_RNG_MT19937, ol.ORA_SAMPLER_REJECTION, 3, 0, 0, cam_r.image_height, 4,
                      img_o, None)
    assert np.array_equal(np.frombuffer(img_r, dtype=np.float64), np.frombuffer(img_o, dtype=np.float64), equal_nan=True)
    oracle.ora_scene_destroy(osc)
    r.ref_scene_free(h)


def test_to_byte_matches_reference(oracle):
    r = ol.ref()
    xs = np.concatenate([np.linspace(-1, 2, 3001), np.random.default_rng(0).random(2000) ** 2])
    assert all(oracle.ora_to_byte(float(x)) == r.ref_to_byte(float(x)) for x in xs)
