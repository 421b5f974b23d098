"""The per-thread device functions of csrc/ (rt_device.h, rt_exact.h, rt_bvh.h, rt_flatten.h) compiled for
the host (tests/emu, a TEST build) and checked against the oracle: scene flattening, the LBVH build logic,
FP64 parity traversal, FP32 traversal and the shading arithmetic.  CPU only.  The GPU tests run the same
checks against the real kernels through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol
from conftest import oracle_segments
from rt_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu")])
    em = C.CDLL(os.path.join(HERE, "emu", "libemu.so"))
    em.emu_scene_create.restype = C.c_void_p
    em.emu_scene_create.argtypes = [C.POINTER(abi.rt_scene_desc)]
    em.emu_scene_destroy.argtypes = [C.c_void_p]
    em.emu_trace.argtypes = [C.c_void_p, C.POINTER(abi.rt_ray), C.c_int64, C.c_int, C.c_uint64, C.POINTER(abi.rt_hit)]
    for f in ("emu_scene_check_bvh", "emu_scene_nodes", "emu_scene_leaves"):
        getattr(em, f).argtypes = [C.c_void_p]
    em.emu_render.argtypes = [C.c_void_p, C.POINTER(abi.rt_camera), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64,
                              C.c_double, C.POINTER(C.c_float), C.POINTER(C.c_uint64)]
    return em


CASES = [("spheres", 11, -1, 96), ("spheres", 40, -1, 64), ("spheres_textured", 12, -1, 64), ("cornell", 0, -1, 48),
         ("cornell_smoke", 0, -1, 48), ("final", 5, 60, 64)]


@pytest.mark.parametrize("name,p0,p1,width", CASES)
def test_flatten_build_trace_shade(oracle, emu, host_scenes, name, p0, p1, width):
    hs = host_scenes(name, p0, p1)
    desc = hs.desc
    cfg = hs.camera_config(width, 4, 8)
    es = emu.emu_scene_create(desc)
    assert es
    n_leaf = desc.contents.n_objects if name != "final" and "cornell" not in name else emu.emu_scene_leaves(es)
    assert emu.emu_scene_leaves(es) == n_leaf
    assert emu.emu_scene_check_bvh(es) == 0  # every leaf once, children inside parents
    assert emu.emu_scene_nodes(es) <= max(1, emu.emu_scene_leaves(es) - 1)
    osc = oracle.ora_scene_create(desc)
    img, rays, cnt = oracle_segments(oracle, osc, cfg, ol.ORA_RNG_PHILOX, ol.ORA_SAMPLER_POLAR, 42, 1)
    n = len(rays)
    want = (abi.rt_hit * n)()
    oracle.ora_trace(osc, rays, n, 1, ol.ORA_RNG_PHILOX, 42, want)
    exact, fast = (abi.rt_hit * n)(), (abi.rt_hit * n)()
    emu.emu_trace(es, rays, n, abi.RT_TRACE_EXACT_F64, 42, exact)
    emu.emu_trace(es, rays, n, abi.RT_TRACE_FAST_F32, 42, fast)
    a, b, c = ol.hits_to_numpy(want), ol.hits_to_numpy(exact), ol.hits_to_numpy(fast)
    assert np.array_equal(a["prim"], b["prim"]) and np.array_equal(a["object"], b["object"])
    assert np.array_equal(a["front_face"], b["front_face"])
    assert np.array_equal(a["t"], b["t"])  # host libm == oracle libm, so even medium hits are bit-exact here
    assert (a["prim"] != c["prim"]).mean() < 2e-3  # FP32: only silhouette / grazing rays may flip
    # FP32 wavefront arithmetic follows the oracle's FP64 paths (same Philox stream)
    cam = abi.rt_camera()
    oracle.ora_camera_init(C.byref(cfg), C.byref(cam))
    npix = cam.image_width * cam.image_height
    out = (C.c_float * (npix * 3))()
    segs = C.c_uint64()
    emu.emu_render(es, C.byref(cam), 0, 4, 2, 8, 42, 0.25, out, C.byref(segs))
    got = np.frombuffer(out, dtype=np.float32).reshape(-1, 3).astype(np.float64)
    follows = np.abs(got - img).max(axis=1) < 2e-3
    assert follows.mean() > 0.97
    assert abs(got.mean() - img.mean()) < 0.01 * img.mean()
    emu.emu_scene_destroy(es)
    oracle.ora_scene_destroy(osc)


def test_degenerate_scenes(emu):
    # empty scene and single-primitive scene build valid one-node trees
    empty = abi.rt_scene_desc()
    es = emu.emu_scene_create(C.byref(empty))
    assert es and emu.emu_scene_leaves(es) == 0 and emu.emu_scene_check_bvh(es) == 0
    ray = (abi.rt_ray * 1)()
    ray[0].direction[:] = (0, 0, -1)
    ray[0].t_min, ray[0].t_max = 0.001, float("inf")
    hit = (abi.rt_hit * 1)()
    for mode in (0, 1):
        emu.emu_trace(es, ray, 1, mode, 0, hit)
        assert hit[0].prim == -1 and hit[0].t == float("inf")
    emu.emu_scene_destroy(es)

    sph = abi.rt_sphere(radius=0.5, material=0, xform=-1, object=0)
    sph.center0[:] = (0, 0, -2)
    mat = abi.rt_material(type=abi.RT_MAT_LAMBERTIAN, texture=0)
    tex = abi.rt_texture(type=abi.RT_TEX_SOLID, even=-1, odd=-1, perlin=-1)
    one = abi.rt_scene_desc(n_spheres=1, n_materials=1, n_textures=1, n_objects=1, spheres=C.pointer(sph),
                            materials=C.pointer(mat), textures=C.pointer(tex))
    es = emu.emu_scene_create(C.byref(one))
    assert es and emu.emu_scene_leaves(es) == 1 and emu.emu_scene_check_bvh(es) == 0
    for mode in (0, 1):
        emu.emu_trace(es, ray, 1, mode, 0, hit)
        assert hit[0].prim == 0 and abs(hit[0].t - 1.5) < 1e-6
    emu.emu_scene_destroy(es)


def test_invalid_scenes_are_rejected(emu):
    sph = abi.rt_sphere(radius=0.5, material=3, xform=-1, object=0)  # material index out of range
    bad = abi.rt_scene_desc(n_spheres=1, n_objects=1, spheres=C.pointer(sph))
    assert not emu.emu_scene_create(C.byref(bad))


def test_image_texture_scene(oracle, emu, host_scenes, tmp_path):
    """Image textures (north-star surface that the reference lacks): sphere (u, v) of Sphere.cpp:136-140 and
    quad (alpha, beta) of Plane.cpp:93-102 index the texel array; FP32 device arithmetic against the oracle."""
    from rt_b200 import host

    hs = host_scenes("earth", 0, -1)
    d = hs.desc.contents
    assert d.n_images == 1 and d.images[0].width == 512 and d.images[0].height == 256
    cfg = hs.camera_config(96, 4, 6)
    es = emu.emu_scene_create(hs.desc)
    assert es and emu.emu_scene_check_bvh(es) == 0
    osc = oracle.ora_scene_create(hs.desc)
    img, rays, cnt = oracle_segments(oracle, osc, cfg, ol.ORA_RNG_PHILOX, ol.ORA_SAMPLER_POLAR, 9, 1)
    cam = abi.rt_camera()
    oracle.ora_camera_init(C.byref(cfg), C.byref(cam))
    npix = cam.image_width * cam.image_height
    out = (C.c_float * (npix * 3))()
    emu.emu_render(es, C.byref(cam), 0, 4, 2, 6, 9, 0.25, out, None)
    got = np.frombuffer(out, dtype=np.float32).reshape(-1, 3).astype(np.float64)
    follows = np.abs(got - img).max(axis=1) < 4e-3  # a texel boundary may flip in FP32
    assert follows.mean() > 0.95
    assert abs(got.mean() - img.mean()) < 0.01 * img.mean()
    # the textured globe really shows the map: green land, blue ocean, white caps are all present
    centre = img.reshape(cam.image_height, cam.image_width, 3)[20:34, 40:56].reshape(-1, 3)
    assert centre[:, 2].max() > 1.5 * centre[:, 0].min() and centre.std() > 0.02
    # JSON round trip carries the image through a PPM file beside the JSON
    path = str(tmp_path / "earth.json")
    hs.save_json(path)
    back = host.HostScene.from_json(path)
    db = back.desc.contents
    assert db.n_images == 1 and (db.images[0].width, db.images[0].height) == (512, 256)
    n = 512 * 256 * 3
    assert C.string_at(db.images[0].rgb, n) == C.string_at(d.images[0].rgb, n)
    assert ol.desc_bytes(db)[5:7] == ol.desc_bytes(d)[5:7]  # materials, textures
    emu.emu_scene_destroy(es)
    oracle.ora_scene_destroy(osc)


def test_fast_division_by_invariant_integers(emu):
    """FastDiv (rt_device.h) maps path ids to pixels in every kernel: exact for every divisor and x < 2^31."""
    emu.emu_fastdiv.restype = C.c_uint32
    emu.emu_fastdiv.argtypes = [C.c_uint32, C.c_uint32]
    rng = np.random.default_rng(7)
    divisors = [1, 2, 3, 5, 7, 8, 10, 64, 225, 400, 1080, 1920, 3840, 2073600, 8294400, 2**30, 2**31 - 1]
    divisors += [int(d) for d in rng.integers(1, 2**31, 200)]
    for d in divisors:
        xs = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2**31 - 1] + [int(x) for x in rng.integers(0, 2**31, 50)]
        xs += [q * d + r for q in (3, 1000, (2**31 - 1) // d) for r in (-1, 0, 1)]
        for x in xs:
            if 0 <= x < 2**31:
                assert emu.emu_fastdiv(d, x) == x // d, (d, x)


@pytest.mark.parametrize("name,p0", [("spheres", 11), ("final", 5), ("cornell", 0)])
def test_all_bvh_builders_answer_alike(oracle, emu, host_scenes, monkeypatch, name, p0):
    """Host SAH tree (rt_sah.h), PLOC and the Karras LBVH (rt_bvh.h), and RT_BVH=best (SAH or PLOC by surface-area sum):
    all consistent, and the FP64 traversal finds the same hits in each; the SAH tree must not need more node visits
    than the LBVH on the bigger scenes, and `best` must visit as few nodes as the better of SAH and PLOC."""
    emu.emu_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
    hs = host_scenes(name, p0, 60 if name == "final" else -1)
    cfg = hs.camera_config(64, 1, 8)
    osc = oracle.ora_scene_create(hs.desc)
    _, rays, _ = oracle_segments(oracle, osc, cfg, ol.ORA_RNG_PHILOX, ol.ORA_SAMPLER_POLAR, 7, 1)
    n = len(rays)
    hits, visits = {}, {}
    for mode in ("lbvh", "sah", "ploc", "best"):
        monkeypatch.setenv("RT_BVH", mode)
        es = emu.emu_scene_create(hs.desc)
        assert es and emu.emu_scene_check_bvh(es) == 0
        out = (abi.rt_hit * n)()
        a, b = C.c_uint64(), C.c_uint64()
        emu.emu_stats(C.byref(a), C.byref(b), 1)
        emu.emu_trace(es, rays, n, abi.RT_TRACE_FAST_F32, 7, out)
        emu.emu_stats(C.byref(a), C.byref(b), 1)
        visits[mode] = a.value
        emu.emu_trace(es, rays, n, abi.RT_TRACE_EXACT_F64, 7, out)
        hits[mode] = ol.hits_to_numpy(out)
        emu.emu_scene_destroy(es)
    for mode in ("sah", "ploc", "best"):
        for k in ("t", "prim", "object", "front_face"):
            assert np.array_equal(hits["lbvh"][k], hits[mode][k]), (mode, k)
    if name != "cornell":  # 13 primitives: either tree is two levels
        assert visits["sah"] < visits["lbvh"], visits
        assert visits["best"] <= 1.02 * min(visits["sah"], visits["ploc"]), visits
    oracle.ora_scene_destroy(osc)


def moved_spheres(desc, rng, count, first=3):
    """A copy of desc's sphere array with `count` surface spheres from `first` on displaced / resized, the
    changed slice, and a scene description that uses the copy."""
    d = desc.contents
    spheres = (abi.rt_sphere * d.n_spheres)()
    C.memmove(spheres, d.spheres, C.sizeof(spheres))
    for i in range(first, first + count):
        s = spheres[i]
        for a in range(3):
            s.center0[a] += float(rng.uniform(-3.0, 3.0)) * (0.2 if a == 1 else 1.0)
        s.center_dir[1] = float(rng.uniform(0.0, 0.5))
        s.radius = float(rng.uniform(0.1, 0.45))
    changed = (abi.rt_sphere * count)(*[spheres[i] for i in range(first, first + count)])
    d2 = abi.rt_scene_desc()
    C.memmove(C.byref(d2), C.byref(d), C.sizeof(d2))
    d2.spheres = C.cast(spheres, C.POINTER(abi.rt_sphere))
    return spheres, changed, d2


def test_update_spheres_refits_the_tree(oracle, emu, host_scenes):
    """rt_scene_update_spheres (rt_scene.cu, mirrored in the emu): moved / resized spheres are re-baked, the BVH4
    is refitted bottom-up, and the updated scene answers exactly like a scene built from the new description."""
    emu.emu_scene_update_spheres.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.rt_sphere)]
    hs = host_scenes("spheres", 11)
    rng = np.random.default_rng(11)
    first, count = 3, 120
    keep, changed, d2 = moved_spheres(hs.desc, rng, count, first)
    es = emu.emu_scene_create(hs.desc)
    assert emu.emu_scene_update_spheres(es, first, count, changed) == 0
    assert emu.emu_scene_check_bvh(es) == 0  # every child box still inside its parent's slot
    fresh = emu.emu_scene_create(C.byref(d2))
    cfg = hs.camera_config(96, 1, 8)
    osc = oracle.ora_scene_create(C.byref(d2))
    _, rays, _ = oracle_segments(oracle, osc, cfg, ol.ORA_RNG_PHILOX, ol.ORA_SAMPLER_POLAR, 5, 1)
    n = len(rays)
    want = (abi.rt_hit * n)()
    oracle.ora_trace(osc, rays, n, 1, ol.ORA_RNG_PHILOX, 5, want)
    a, b = (abi.rt_hit * n)(), (abi.rt_hit * n)()
    emu.emu_trace(es, rays, n, abi.RT_TRACE_EXACT_F64, 5, a)
    emu.emu_trace(fresh, rays, n, abi.RT_TRACE_EXACT_F64, 5, b)
    w, ha, hb = ol.hits_to_numpy(want), ol.hits_to_numpy(a), ol.hits_to_numpy(b)
    for k in ("t", "prim", "object", "front_face"):
        assert np.array_equal(ha[k], hb[k]), k
        assert np.array_equal(ha[k], w[k]), k
    moved_hit = np.isin(w["prim"], np.arange(first, first + count)).sum()
    assert moved_hit > 50  # the rays do see the moved spheres
    # the FP32 traversal over the refitted tree finds them as well
    c = (abi.rt_hit * n)()
    emu.emu_trace(es, rays, n, abi.RT_TRACE_FAST_F32, 5, c)
    assert (ol.hits_to_numpy(c)["prim"] != w["prim"]).mean() < 2e-3
    # a boundary sphere of a medium has no leaf of its own
    smoke = host_scenes("final", 5, 60)
    d = smoke.desc.contents
    boundary = [i for i in range(d.n_spheres) if d.spheres[i].flags & abi.RT_PRIM_BOUNDARY]
    if boundary:
        es2 = emu.emu_scene_create(smoke.desc)
        one = (abi.rt_sphere * 1)(d.spheres[boundary[0]])
        assert emu.emu_scene_update_spheres(es2, boundary[0], 1, one) != 0
        emu.emu_scene_destroy(es2)
    emu.emu_scene_destroy(es)
    emu.emu_scene_destroy(fresh)
    oracle.ora_scene_destroy(osc)
    del keep


def test_update_quads_refits_the_tree(oracle, emu, host_scenes):
    """rt_scene_update_quads on the Cornell box: the two boxes' quads and the light move / resize; the refitted
    scene answers exactly like a scene built from the new description and like the oracle on it."""
    emu.emu_scene_update_quads.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(abi.rt_quad)]
    hs = host_scenes("cornell", 0)
    d = hs.desc.contents
    rng = np.random.default_rng(2)
    quads = (abi.rt_quad * d.n_quads)()
    C.memmove(quads, d.quads, C.sizeof(quads))
    first, count = 5, d.n_quads - 5  # everything but the five walls
    for i in range(first, first + count):
        q = quads[i]
        for a in range(3):
            q.corner[a] += float(rng.uniform(-40.0, 40.0))
            q.u[a] *= float(rng.uniform(0.7, 1.2))
            q.v[a] *= float(rng.uniform(0.7, 1.2))
    changed = (abi.rt_quad * count)(*[quads[i] for i in range(first, first + count)])
    d2 = abi.rt_scene_desc()
    C.memmove(C.byref(d2), C.byref(d), C.sizeof(d2))
    d2.quads = C.cast(quads, C.POINTER(abi.rt_quad))
    es = emu.emu_scene_create(hs.desc)
    assert emu.emu_scene_update_quads(es, first, count, changed) == 0
    assert emu.emu_scene_check_bvh(es) == 0
    fresh = emu.emu_scene_create(C.byref(d2))
    cfg = hs.camera_config(64, 1, 6)
    osc = oracle.ora_scene_create(C.byref(d2))
    _, rays, _ = oracle_segments(oracle, osc, cfg, ol.ORA_RNG_PHILOX, ol.ORA_SAMPLER_POLAR, 4, 1)
    n = len(rays)
    want = (abi.rt_hit * n)()
    oracle.ora_trace(osc, rays, n, 1, ol.ORA_RNG_PHILOX, 4, want)
    a, b = (abi.rt_hit * n)(), (abi.rt_hit * n)()
    emu.emu_trace(es, rays, n, abi.RT_TRACE_EXACT_F64, 4, a)
    emu.emu_trace(fresh, rays, n, abi.RT_TRACE_EXACT_F64, 4, b)
    w, ha, hb = ol.hits_to_numpy(want), ol.hits_to_numpy(a), ol.hits_to_numpy(b)
    for k in ("t", "prim", "object", "front_face"):
        assert np.array_equal(ha[k], hb[k]), k
        assert np.array_equal(ha[k], w[k]), k
    assert np.isin(w["prim"], d.n_spheres + np.arange(first, first + count)).sum() > 50
    emu.emu_scene_destroy(es)
    emu.emu_scene_destroy(fresh)
    oracle.ora_scene_destroy(osc)


@pytest.mark.parametrize("width,height,n_ranks,tile_rows", [(1920, 1080, 1, 8), (1920, 1080, 3, 8), (400, 225, 1, 8),
                                                            (64, 48, 2, 4), (40, 40, 4, 8), (33, 21, 2, 8), (8, 4, 1, 4)])
def test_path_numbering_is_a_bijection(emu, width, height, n_ranks, tile_rows):
    """PathMap / path_to_pixel (rt_device.h): every owned pixel gets exactly one path per sample, the film index is
    row-major over the owned scanlines, scanlines follow the tile ownership rule, and when the 8 x 4 block order is
    in use 32 consecutive paths cover one 8 x 4 pixel block."""
    from rt_b200 import distributed

    emu.emu_path_map.argtypes = [C.c_int] * 6 + [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
    for rank in range(n_ranks):
        rows = distributed.owned_rows(height, rank, n_ranks, tile_rows)
        n_owned = len(rows) * width
        for allow in (1, 0):
            n_samples = 2
            out = (C.c_uint32 * (4 * n_owned * n_samples))()
            tiled = emu.emu_path_map(width, height, rank, n_ranks, tile_rows, allow, 0, n_owned * n_samples, out)
            m = np.frombuffer(out, dtype=np.uint32).reshape(-1, 4).astype(np.int64)
            assert tiled == int(bool(allow) and width % 8 == 0 and len(rows) % 4 == 0 and tile_rows % 4 == 0)
            for s in range(n_samples):
                part = m[s * n_owned:(s + 1) * n_owned]
                assert (part[:, 0] == s).all()
                assert np.array_equal(np.sort(part[:, 1]), np.arange(n_owned))       # every film slot once
                local_row, col = part[:, 1] // width, part[:, 1] % width
                assert np.array_equal(col, part[:, 3])
                assert np.array_equal(np.asarray(rows)[local_row], part[:, 2])       # tile ownership rule
                if tiled:
                    blocks = part.reshape(-1, 32, 4)
                    assert ((blocks[:, :, 3].max(axis=1) - blocks[:, :, 3].min(axis=1)) == 7).all()
                    assert ((local_row.reshape(-1, 32).max(axis=1) - local_row.reshape(-1, 32).min(axis=1)) == 3).all()
                else:
                    assert np.array_equal(part[:, 1], np.arange(n_owned))            # scanline order


def test_fp32_slab_test_never_rejects_a_box_the_exact_test_enters(emu):
    """node_visit's slab test (FP32, fused multiply-adds, precomputed o/d) may enter boxes the exact test would
    skip, never the other way round - otherwise the fast traversal could lose hits.  The guarantee of the pipeline:
    every box stored in a node is the geometry's box moved two ulps outwards (to_build_box), and the FP32 test on the
    STORED box enters whenever the slab test in exact rational arithmetic enters the geometry's box - over magnitudes
    from 1e-3 to 1e6, axis-parallel and grazing rays, tiny direction components, origins inside, on and far outside
    the box.  (No absolute allowance: see RayTrav in rt_device.h for the argument.)"""
    from fractions import Fraction

    emu.emu_box_test.argtypes = [C.POINTER(C.c_float)] * 4 + [C.c_float, C.c_float]
    emu.emu_box_test_biased.argtypes = [C.POINTER(C.c_float)] * 4 + [C.c_float, C.c_float, C.POINTER(C.c_int)]
    rng = np.random.default_rng(17)

    def exact_accepts(lo, hi, o, d, tmin, tmax):
        t0, t1 = Fraction(float(tmin)), Fraction(float(tmax)) if np.isfinite(tmax) else None
        for a in range(3):
            if d[a] == 0:
                if not (lo[a] <= o[a] <= hi[a]):
                    return False
                continue
            ta = (Fraction(float(lo[a])) - Fraction(float(o[a]))) / Fraction(float(d[a]))
            tb = (Fraction(float(hi[a])) - Fraction(float(o[a]))) / Fraction(float(d[a]))
            near, far = min(ta, tb), max(ta, tb)
            t0 = max(t0, near)
            t1 = far if t1 is None else min(t1, far)
        return t1 is None or t0 <= t1

    rejected_but_exact = 0
    entered = exact = 0
    for trial in range(24000):
        scale = np.float32(10.0 ** rng.uniform(-3, 6))
        centre = (rng.uniform(-1, 1, 3) * scale).astype(np.float32)
        half = (rng.uniform(1e-3, 1, 3) * scale * 10.0 ** rng.uniform(-3, 0)).astype(np.float32)
        lo, hi = (centre - half).astype(np.float32), (centre + half).astype(np.float32)
        kind = trial % 6
        if kind == 0:    # origin far away, aimed at a point of the box surface (grazing / corner hits)
            target = np.where(rng.random(3) < 0.5, lo, hi).astype(np.float32)
            o = (centre + rng.normal(size=3) * scale * 10.0 ** rng.uniform(0, 3)).astype(np.float32)
            d = (target.astype(np.float64) - o.astype(np.float64)).astype(np.float32)
        elif kind == 1:  # axis-parallel ray along an edge plane
            o = (centre + rng.normal(size=3) * scale * 3).astype(np.float32)
            d = np.zeros(3, dtype=np.float32)
            a = int(rng.integers(0, 3))
            d[a] = np.float32(rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(-3, 3))
            b = (a + 1) % 3
            o[b] = lo[b] if rng.random() < 0.5 else hi[b]
        elif kind == 2:  # origin inside the box
            o = (lo + (hi - lo) * rng.random(3)).astype(np.float32)
            d = rng.normal(size=3).astype(np.float32)
        elif kind == 3:  # tiny direction components
            o = (centre + rng.normal(size=3) * scale * 5).astype(np.float32)
            d = (rng.normal(size=3) * 10.0 ** rng.uniform(-12, 0, 3)).astype(np.float32)
        elif kind == 4:  # skimming: aimed into the box, but on one axis the origin sits within a few ulps of a face
            a = int(rng.integers(0, 3))  # plane and moves along that axis by next to nothing (|o / d| astronomically large)
            o = (centre + rng.normal(size=3) * scale * 10.0 ** rng.uniform(0, 2)).astype(np.float32)
            d = ((lo + (hi - lo) * rng.random(3)).astype(np.float64) - o.astype(np.float64)).astype(np.float32)
            face = lo[a] if rng.random() < 0.5 else hi[a]
            for _ in range(int(rng.integers(0, 4))):
                face = np.nextafter(face, np.float32(rng.choice([-np.inf, np.inf])))
            o[a] = face
            d[a] = np.float32(rng.choice([-1.0, 1.0]) * 10.0 ** rng.uniform(-38, -3)) if rng.random() < 0.8 else np.float32(0.0)
        else:            # generic
            o = (centre + rng.normal(size=3) * scale * 5).astype(np.float32)
            d = (centre.astype(np.float64) + rng.normal(size=3) * half * 1.5 - o).astype(np.float32)
        if not np.any(d != 0):
            continue
        tmin, tmax = np.float32(0.001), np.float32(np.inf if trial % 3 else 10.0 ** rng.uniform(-2, 7))
        lo_s, hi_s = lo.copy(), hi.copy()  # the stored box: two ulps outwards, as to_build_box makes it
        for _ in range(2):
            lo_s, hi_s = np.nextafter(lo_s, np.float32(-np.inf)), np.nextafter(hi_s, np.float32(np.inf))
        fast = emu.emu_box_test(lo_s.ctypes.data_as(C.POINTER(C.c_float)), hi_s.ctypes.data_as(C.POINTER(C.c_float)),
                                o.ctypes.data_as(C.POINTER(C.c_float)), d.ctypes.data_as(C.POINTER(C.c_float)), tmin, tmax)
        # the device's one-instruction reciprocal may be an ulp off the correctly rounded 1 / d, either way, per axis
        bias = (C.c_int * 3)(*[int(b) for b in rng.integers(-1, 2, 3)])
        fast_biased = emu.emu_box_test_biased(lo_s.ctypes.data_as(C.POINTER(C.c_float)), hi_s.ctypes.data_as(C.POINTER(C.c_float)),
                                              o.ctypes.data_as(C.POINTER(C.c_float)), d.ctypes.data_as(C.POINTER(C.c_float)),
                                              tmin, tmax, bias)
        want = exact_accepts(lo, hi, o, d, tmin, tmax)
        entered += fast
        exact += want
        if want and not (fast and fast_biased):
            rejected_but_exact += 1
    assert rejected_but_exact == 0
    assert exact > 1500 and entered >= exact  # the cases do exercise both outcomes


def test_fp32_sphere_test_stays_accurate_far_from_the_sphere(emu):
    """sphere_hit evaluates the discriminant as a (r^2 - |oc - (h/a) d|^2): no cancellation between |oc|^2 and r^2,
    so a radius-1000 ground sphere seen from its surface and a unit sphere seen from 10^4 radii away are both
    intersected at the distance high-precision arithmetic gives (relative error of t below 1e-4 whenever the ray is
    not grazing), where the textbook h^2 - a c form loses every digit in FP32."""
    from decimal import Decimal, getcontext

    getcontext().prec = 60
    emu.emu_sphere_test.argtypes = [C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.c_float, C.c_float, C.POINTER(C.c_float)]
    rng = np.random.default_rng(29)
    fp = lambda v: v.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    checked = textbook_bad = 0
    for trial in range(3000):
        radius = np.float32(10.0 ** rng.uniform(-2, 3))
        centre = (rng.normal(size=3) * 10.0 ** rng.uniform(0, 3)).astype(np.float32)
        # aim at a point well inside the silhouette from `dist` radii away
        dist = 10.0 ** rng.uniform(0.01, 4)
        direction = rng.normal(size=3)
        direction /= np.linalg.norm(direction)
        o = (centre.astype(np.float64) - direction * float(radius) * dist).astype(np.float32)
        off = rng.normal(size=3)
        off -= off.dot(direction) * direction
        off *= float(radius) * rng.uniform(0, 0.8) / max(np.linalg.norm(off), 1e-30)
        d = ((centre.astype(np.float64) + off - o) * 10.0 ** rng.uniform(-2, 2)).astype(np.float32)
        D = lambda x: Decimal(float(x))  # noqa: E731
        oc = [D(centre[k]) - D(o[k]) for k in range(3)]
        a = sum(D(d[k]) * D(d[k]) for k in range(3))
        h = sum(D(d[k]) * oc[k] for k in range(3))
        c = sum(x * x for x in oc) - D(radius) * D(radius)
        disc = h * h - a * c
        if disc <= 0 or c <= 0:  # grazing, or the origin is inside the sphere
            continue
        t_exact = (h - disc.sqrt()) / a
        if t_exact <= Decimal("0.002"):
            continue
        t = C.c_float()
        hit = emu.emu_sphere_test(fp(centre), radius, fp(o), fp(d), np.float32(0.001), np.float32(np.inf), C.byref(t))
        assert hit == 1, (trial, float(radius), dist)
        rel = abs(Decimal(t.value) - t_exact) / t_exact
        assert rel < Decimal("1e-4"), (trial, float(radius), dist, float(rel))
        checked += 1
        # the textbook FP32 discriminant for comparison
        oc32 = (centre - o).astype(np.float32)
        a32, h32 = np.float32(d @ d), np.float32(d @ oc32)
        c32 = np.float32(np.float32(oc32 @ oc32) - radius * radius)
        disc32 = np.float32(h32 * h32 - a32 * c32)
        if disc32 < 0 or abs(Decimal(float((h32 - np.sqrt(max(disc32, np.float32(0)))) / a32)) - t_exact) / t_exact > Decimal("1e-4"):
            textbook_bad += 1
    assert checked > 2000
    assert textbook_bad > 50  # the formulation matters: the plain FP32 form fails on a visible fraction of these rays


def test_fp32_quad_test_against_exact_arithmetic(emu):
    """quad_hit on records baked by rt_flatten.h: rays aimed at points well inside a quad hit it at the distance
    rational arithmetic gives, rays aimed well outside miss, for quads of size 1e-2 .. 1e3 placed up to 1e3 from
    the origin and seen from up to 1e3 sizes away."""
    from fractions import Fraction

    emu.emu_quad_test.argtypes = [C.POINTER(C.c_double)] * 3 + [C.POINTER(C.c_float)] * 2 + [C.c_float, C.c_float,
                                                                                             C.POINTER(C.c_float)]
    rng = np.random.default_rng(31)
    dp = lambda v: v.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    fp = lambda v: v.ctypes.data_as(C.POINTER(C.c_float))  # noqa: E731
    inside = outside = 0
    for trial in range(3000):
        size = 10.0 ** rng.uniform(-2, 3)
        Q = rng.normal(size=3) * 10.0 ** rng.uniform(0, 3)
        u = rng.normal(size=3)
        v = rng.normal(size=3)
        v -= 0.5 * v.dot(u) / u.dot(u) * u  # not degenerate, not necessarily orthogonal
        u *= size / np.linalg.norm(u)
        v *= size * rng.uniform(0.3, 1.0) / np.linalg.norm(v)
        n = np.cross(u, v)
        n /= np.linalg.norm(n)
        want_inside = trial % 2 == 0
        if want_inside:
            al, be = rng.uniform(0.05, 0.95, 2)
        else:
            al, be = rng.uniform(1.05, 3.0) * rng.choice([-1, 1]) + 0.5, rng.uniform(-1, 2)
        target = Q + al * u + be * v
        away = n * rng.choice([-1, 1]) + 0.7 * rng.normal(size=3)
        o = (target + away / np.linalg.norm(away) * size * 10.0 ** rng.uniform(-1, 3)).astype(np.float32)
        d = ((target - o.astype(np.float64)) * 10.0 ** rng.uniform(-1, 1)).astype(np.float32)
        t = C.c_float()
        hit = emu.emu_quad_test(dp(Q), dp(u), dp(v), fp(o), fp(d), np.float32(0.001), np.float32(np.inf), C.byref(t))
        # exact plane distance of the FP32 ray against the FP64 quad
        F = lambda x: Fraction(float(x))  # noqa: E731
        nn = [F(x) for x in np.cross(u, v)]
        denom = sum(nn[k] * F(d[k]) for k in range(3))
        t_exact = sum(nn[k] * (F(Q[k]) - F(o[k])) for k in range(3)) / denom
        if want_inside:
            assert hit == 1, (trial, size)
            # FP32 coordinates of magnitude m carry m * 2^-24 of position error: that, or 1e-4 of the distance
            m = max(np.abs(Q).max() + size, np.abs(o).max())
            dn = d.astype(np.float64)
            cos = abs(dn.dot(n)) / np.linalg.norm(dn)  # grazing incidence stretches a position error along the ray
            allowed = max(t_exact / 10000, Fraction(float(32 * 2.0 ** -24 * m / (np.linalg.norm(dn) * cos))))
            assert abs(Fraction(t.value) - t_exact) < allowed, (trial, size)
            inside += 1
        else:
            assert hit == 0, (trial, size, al, be)
            outside += 1
    assert inside == 1500 and outside == 1500


def test_axis_parallel_rays_do_not_walk_the_whole_tree(emu, host_scenes):
    """A direction with exactly zero components (unit_vector_polar returns them with probability ~2^-22 per draw) makes
    o / d infinite on that axis.  The slab test has no absolute allowance and make_trav keeps |d| >= 1e-20 for the
    box tests, so such a ray sees the boxes an almost-parallel ray sees - it once entered EVERY box (1645 nodes, 3409
    primitive tests on this scene, 2 ms of one GPU lane) - and still finds the FP64 traversal's primitive."""
    emu.emu_stats.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_int]
    hs = host_scenes("final", 20, 1000)
    es = emu.emu_scene_create(hs.desc)
    n_nodes = emu.emu_scene_nodes(es)
    assert n_nodes > 1000
    cases = [((478, 278, -600), (0, 0, 1)), ((478, 278, -600), (0, 0.6, 0.8)), ((478, 278, -600), (1e-30, 0, 1)),
             ((478, 278, -600), (-0.0, 0.001, 1)), ((3000, 300, 200), (-1, 0.0, 0.0)), ((278, 554, 279.5), (0, -1, 0)),
             ((278, 1000, 279.5), (0, -1, 0)), ((123.4, 2000, 77.7), (0, -1, 0)), ((0.5, 0.25, -900), (0, 0, 1))]
    for o, d in cases:
        ray = (abi.rt_ray * 1)()
        for k in range(3):
            ray[0].origin[k], ray[0].direction[k] = o[k], d[k]
        ray[0].time, ray[0].t_min, ray[0].t_max = 0.5, 0.001, float("inf")
        fast, exact = (abi.rt_hit * 1)(), (abi.rt_hit * 1)()
        a, b = C.c_uint64(), C.c_uint64()
        emu.emu_stats(C.byref(a), C.byref(b), 1)
        emu.emu_trace(es, ray, 1, abi.RT_TRACE_FAST_F32, 1, fast)
        emu.emu_stats(C.byref(a), C.byref(b), 1)
        emu.emu_trace(es, ray, 1, abi.RT_TRACE_EXACT_F64, 1, exact)
        assert a.value <= 40 and b.value <= 20, (o, d, a.value, b.value)
        assert (fast[0].prim >= 0) == (exact[0].prim >= 0), (o, d)
        if exact[0].prim >= 0:  # coincident faces of neighbouring floor boxes tie: the distance decides
            assert abs(fast[0].t - exact[0].t) <= 1e-5 * exact[0].t, (o, d)
    emu.emu_scene_destroy(es)
