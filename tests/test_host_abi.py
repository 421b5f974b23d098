"""Host-side logic and the C-ABI surface.  CPU only: no compute entry point is called without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from rt_b200 import abi, distributed, host

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions(path):
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(rth?_[a-z0-9_]+)\s*\(", text))


def exported(path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", path], text=True)
    return {line.split()[-1] for line in out.splitlines() if " T " in line}


def test_cuda_library_exports_every_declared_symbol():
    declared = header_functions(os.path.join(REPO, "include", "rt_b200.h"))
    assert declared == set(abi.RT_B200_SYMBOLS), "abi.py must mirror include/rt_b200.h"
    assert declared <= exported(abi.LIB_PATH)
    lib = abi.load_library()
    assert lib.rt_abi_version() == abi.RT_B200_ABI_VERSION == 2


def test_host_library_exports_every_declared_symbol():
    declared = header_functions(os.path.join(REPO, "include", "rt_host.h"))
    assert declared <= exported(abi.HOST_LIB_PATH)


def test_struct_sizes_match_the_header():
    # sizes the C compiler gives the same declarations (compiled once from the header)
    src = '#include <stdio.h>\n#include "rt_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu",' \
          "sizeof(rt_sphere),sizeof(rt_quad),sizeof(rt_xform_op),sizeof(rt_xform),sizeof(rt_medium),sizeof(rt_material)," \
          "sizeof(rt_texture),sizeof(rt_perlin),sizeof(rt_light),sizeof(rt_scene_desc),sizeof(rt_camera_config)," \
          "sizeof(rt_camera),sizeof(rt_ray),sizeof(rt_hit),sizeof(rt_counters),sizeof(rt_audit),sizeof(rt_audit_sample));return 0;}"
    exe = "/tmp/rt_sizes_test"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(REPO, "include"), "-o", exe], input=src, text=True, check=True)
    sizes = list(map(int, subprocess.check_output([exe], text=True).split()))
    mine = [C.sizeof(t) for t in (abi.rt_sphere, abi.rt_quad, abi.rt_xform_op, abi.rt_xform, abi.rt_medium,
                                  abi.rt_material, abi.rt_texture, abi.rt_perlin, abi.rt_light, abi.rt_scene_desc,
                                  abi.rt_camera_config, abi.rt_camera, abi.rt_ray, abi.rt_hit, abi.rt_counters,
                                  abi.rt_audit, abi.rt_audit_sample)]
    assert sizes == mine


def test_no_cpu_fallback_without_a_device():
    lib = abi.load_library()
    if lib.rt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.rt_context_create(0, C.byref(h)) == abi.RT_ERR_NO_DEVICE
    assert not h.value
    assert b"no CPU fallback" in lib.rt_last_error()


def test_camera_init_rejects_bad_configs():
    lib = abi.load_library()
    cfg = abi.rt_camera_config(image_width=0, samples_per_pixel=1, max_depth=1, aspect_ratio=1.0)
    cam = abi.rt_camera()
    assert lib.rt_camera_init(C.byref(cfg), C.byref(cam)) == abi.RT_ERR_INVALID
    assert lib.rt_camera_init(None, C.byref(cam)) == abi.RT_ERR_INVALID


def test_image_height_rule():
    # Camera::initialize: height = int(width / aspect), at least 1 (Camera.cpp:32-33)
    lib = abi.load_library()
    for w, aspect, want in [(400, 16 / 9, 225), (1920, 16 / 9, 1080), (600, 1.0, 600), (3, 16.0, 1)]:
        cfg = abi.rt_camera_config(image_width=w, samples_per_pixel=1, max_depth=1, aspect_ratio=aspect, vfov=20,
                                   focus_dist=10)
        cfg.lookfrom[:] = (13, 2, 3)
        cfg.vup[:] = (0, 1, 0)
        cam = abi.rt_camera()
        assert lib.rt_camera_init(C.byref(cfg), C.byref(cam)) == 0
        assert (cam.image_width, cam.image_height) == (w, want)


@pytest.mark.parametrize("height,n_ranks,tile_rows", [(1080, 1, 8), (1080, 2, 8), (1080, 8, 8), (225, 3, 16), (7, 4, 2),
                                                      (5, 8, 1), (2160, 8, 32)])
def test_tile_ownership_partitions_the_image(height, n_ranks, tile_rows):
    lib = abi.load_library()
    width = 37
    seen = np.zeros(height, dtype=int)
    for r in range(n_ranks):
        rows = distributed.owned_rows(height, r, n_ranks, tile_rows)
        seen[rows] += 1
        assert lib.rt_film_owned_pixels_for(width, height, r, n_ranks, tile_rows) == len(rows) * width
    assert np.all(seen == 1)
    assert lib.rt_film_owned_pixels_for(width, height, n_ranks, n_ranks, tile_rows) == -1  # rank out of range


def test_assemble_host_inverts_the_partition():
    width, height, tile_rows, n_ranks = 11, 53, 4, 3
    full = np.arange(width * height * 4, dtype=np.float32).reshape(height, width, 4)
    parts = [full[distributed.owned_rows(height, r, n_ranks, tile_rows)].reshape(-1, 4) for r in range(n_ranks)]
    assert np.array_equal(distributed.assemble_host(parts, width, height, tile_rows), full)


def test_ppm_writer(tmp_path):
    # "P3\nW H\n255\n" then one "r g b" line per pixel (StaticCamera.cpp:57, ColorUtility.hpp:30-37)
    img = (np.arange(4 * 3 * 3) * 7 % 256).astype(np.uint8)
    path = tmp_path / "out.ppm"
    host.write_ppm_p3(str(path), 4, 3, img)
    lines = path.read_text().split("\n")
    assert lines[:3] == ["P3", "4 3", "255"]
    body = [list(map(int, l.split())) for l in lines[3:] if l]
    assert np.array_equal(np.array(body).reshape(-1), img)


def test_unknown_builtin_scene_is_an_error():
    with pytest.raises(abi.RtError):
        host.HostScene.builtin("no-such-scene")


class rth_cli_options(C.Structure):
    _fields_ = [("width", C.c_int), ("samples", C.c_int), ("depth", C.c_int), ("camera_dynamic", C.c_int),
                ("use_parallelism", C.c_int), ("use_bvh", C.c_int), ("use_gpu", C.c_int), ("debug", C.c_int),
                ("help", C.c_int), ("output", C.c_char * 256), ("scene", C.c_char * 256), ("seed", C.c_uint64),
                ("gpus", C.c_int), ("frames", C.c_int), ("keys", C.c_char * 256), ("headless", C.c_int), ("adaptive", C.c_int)]


def parse_cli(args):
    lib = host.load_host_library()
    argv = (C.c_char_p * (len(args) + 1))(b"raytracer", *[a.encode() for a in args])
    opt = rth_cli_options()
    lib.rth_cli_parse.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(rth_cli_options)]
    rc = lib.rth_cli_parse(len(args) + 1, argv, C.byref(opt))
    return rc, opt


def test_cli_defaults_and_flags_match_the_reference():
    # defaults of CLIOptions (input/CLI.hpp:8-27)
    rc, o = parse_cli([])
    assert rc == 0 and (o.width, o.samples, o.depth) == (600, 100, 50)
    assert o.camera_dynamic == 0 and o.output == b"image.ppm" and not (o.use_parallelism or o.use_bvh or o.use_gpu or o.debug)
    rc, o = parse_cli(["--camera", "dynamic", "-p", "-b", "-g", "-d", "--width", "1920", "--samples", "16", "--depth", "8",
                       "--output", "x.ppm", "--scene", "spheres", "--seed", "7", "--gpus", "2", "--keys", "..w.a"])
    assert rc == 0 and o.camera_dynamic == 1 and (o.width, o.samples, o.depth) == (1920, 16, 8)
    assert o.use_parallelism and o.use_bvh and o.use_gpu and o.debug and o.output == b"x.ppm"
    assert o.scene == b"spheres" and o.seed == 7 and o.gpus == 2 and o.keys == b"..w.a"
    # the reference's error cases (CLI.cpp:13-92)
    for bad in (["--camera", "orbit"], ["--camera"], ["--width"], ["--width", "abc"], ["--bogus"], ["--samples", "0"]):
        rc, _ = parse_cli(bad)
        assert rc != 0, bad
    rc, o = parse_cli(["-h"])
    assert rc == 0 and o.help == 1


@pytest.mark.parametrize("name,p0,p1", [("spheres", 4, -1), ("spheres_textured", 5, -1), ("cornell", 0, -1),
                                        ("cornell_smoke", 0, -1), ("final", 3, 20)])
def test_json_scene_round_trip(tmp_path, name, p0, p1):
    import oracle_lib as ol

    a = host.HostScene.builtin(name, 99, p0, p1)
    path = str(tmp_path / f"{name}.json")
    a.save_json(path)
    b = host.HostScene.from_json(path)
    da, db = a.desc.contents, b.desc.contents
    pa, pb = ol.desc_bytes(da), ol.desc_bytes(db)
    # materials, textures, perlin tables, media and lights survive bit for bit; primitives too, except that
    # every object gets its own copy of a shared instance chain (index fields differ, geometry does not)
    names = ["spheres", "quads", "xform_ops", "xforms", "media", "materials", "textures", "perlins", "lights"]
    for nm, x, y in zip(names, pa, pb):
        if nm in ("materials", "textures", "perlins", "media", "lights"):
            assert x == y, nm
    assert (da.n_spheres, da.n_quads, da.n_objects) == (db.n_spheres, db.n_quads, db.n_objects)
    sa = np.frombuffer(pa[0], dtype=np.uint8).reshape(da.n_spheres, -1) if da.n_spheres else None
    sb = np.frombuffer(pb[0], dtype=np.uint8).reshape(db.n_spheres, -1) if db.n_spheres else None
    if sa is not None:
        assert np.array_equal(sa[:, :60], sb[:, :60])  # centres, motion, radius, material
    ca, cb = a.camera_config(100, 4, 5), b.camera_config(100, 4, 5)
    assert bytes(ca) == bytes(cb)
    # same closest hits through the oracle
    o = ol.oracle()
    n = 40 * 22 if ca.aspect_ratio > 1.5 else 40 * 40
    cfg = a.camera_config(40, 1, 4)
    rays = (abi.rt_ray * n)()
    o.ora_primary_rays(C.byref(cfg), ol.ORA_RNG_PHILOX, ol.ORA_SAMPLER_POLAR, 3, 0, 0, rays)
    sca, scb = o.ora_scene_create(a.desc), o.ora_scene_create(b.desc)
    ha, hb = (abi.rt_hit * n)(), (abi.rt_hit * n)()
    o.ora_trace(sca, rays, n, 1, ol.ORA_RNG_PHILOX, 3, ha)
    o.ora_trace(scb, rays, n, 1, ol.ORA_RNG_PHILOX, 3, hb)
    assert bytes(ha) == bytes(hb)
    o.ora_scene_destroy(sca)
    o.ora_scene_destroy(scb)


def test_json_errors_are_reported(tmp_path):
    p = tmp_path / "bad.json"
    p.write_text('{"camera": {"aspect_ratio": 1.0}, "world": []}')
    with pytest.raises(abi.RtError):
        host.HostScene.from_json(str(p))
    with pytest.raises(abi.RtError):
        host.HostScene.from_json(str(tmp_path / "missing.json"))


def test_damaged_json_scenes_never_crash_the_loader(tmp_path):
    """Truncations and single-character corruptions of a valid scene file: rth_scene_load_json either loads a
    scene or reports an error.  Runs in a child process so that a crash of the parser would be seen as one."""
    import subprocess
    import sys

    hs = host.HostScene.builtin("final", 5, 2, 6)
    good = tmp_path / "good.json"
    hs.save_json(str(good))
    text = good.read_text()
    rng = np.random.default_rng(5)
    cases = []
    for cut in sorted(set(int(x) for x in rng.integers(1, len(text), 60))):
        cases.append(text[:cut])
    for pos in rng.integers(0, len(text), 120):
        ch = "{}[],:\"0-9.eE xyz\n"[int(rng.integers(0, 18))]
        cases.append(text[:int(pos)] + ch + text[int(pos) + 1:])
    cases += ["", "{", "[]", "null", '{"world": 3}', '{"world": [{"type": "Sphere"}]}', '{"world": [{"type": "Nope"}], "camera": {}}',
              '{"camera": {"image_width": 1e99}, "world": []}', "{" * 5000, "[" * 5000 + "]" * 5000,
              '{"world":' + "[" * 2000000 + "]" * 2000000 + "}"]  # would overflow the stack of an unbounded recursive parser
    for i, c in enumerate(cases):
        (tmp_path / f"case{i}.json").write_text(c)
    child = f"""
import sys
sys.path.insert(0, {os.path.dirname(host.__file__)!r} + "/..")
from rt_b200 import abi, host
ok = bad = 0
for i in range({len(cases)}):
    try:
        host.HostScene.from_json({str(tmp_path)!r} + "/case%d.json" % i).close()
        ok += 1
    except abi.RtError:
        bad += 1
print("loaded", ok, "rejected", bad)
"""
    r = subprocess.run([sys.executable, "-c", child], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    ok, bad = int(r.stdout.split()[1]), int(r.stdout.split()[3])
    assert ok + bad == len(cases) and bad > len(cases) // 3


def test_image_texture_files_are_validated(tmp_path):
    """Image textures come from PPM files beside the JSON: P6 and P3 load, damaged files are refused."""
    hs = host.HostScene.builtin("earth", 1)
    good = tmp_path / "earth.json"
    hs.save_json(str(good))
    ppm = tmp_path / "earth.json.image0.ppm"
    assert ppm.exists()
    data = ppm.read_bytes()
    header_end = data.index(b"255\n") + 4
    w, h = [int(x) for x in data[:header_end].split()[1:3]]
    a = host.HostScene.from_json(str(good))
    assert a.desc.contents.n_images == 1 and a.desc.contents.images[0].width == w
    # the same pixels as ASCII P3 with a comment line
    px = np.frombuffer(data[header_end:], dtype=np.uint8)
    ppm.write_text("P3\n# comment\n%d %d\n255\n" % (w, h) + " ".join(str(int(v)) for v in px) + "\n")
    b = host.HostScene.from_json(str(good))
    n = w * h * 3
    assert bytes(C.cast(b.desc.contents.images[0].rgb, C.POINTER(C.c_uint8 * n)).contents) == px.tobytes()
    for bad in (data[:header_end + 100], b"P5\n4 4\n255\n" + bytes(16), b"P6\n4 4\n65535\n" + bytes(96),
                b"P6\n2000000000 2000000000\n255\n", b"P6\n-4 4\n255\n" + bytes(48), b"P3\n2 2\n255\n1 2 3 999", b""):
        ppm.write_bytes(bad)
        with pytest.raises(abi.RtError):
            host.HostScene.from_json(str(good))


def build_c_example(tmp_path):
    """examples/render_ppm.c: both headers from plain C11 (-pedantic), linked against the two shared libraries."""
    exe = str(tmp_path / "render_ppm")
    host_dir = os.path.dirname(abi.HOST_LIB_PATH)
    lib_dir = os.path.dirname(abi.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(REPO, "include"),
                           os.path.join(REPO, "examples", "render_ppm.c"), "-L", host_dir, "-lrt_host", "-L", lib_dir, "-lrt_b200",
                           "-lm", f"-Wl,-rpath,{host_dir}", f"-Wl,-rpath,{lib_dir}", "-o", exe])
    return exe


def test_c_example_builds_and_refuses_to_render_on_the_cpu(tmp_path):
    exe = build_c_example(tmp_path)
    lib = abi.load_library()
    r = subprocess.run([exe, "cornell", "48", "4", "4", str(tmp_path / "c.ppm")], capture_output=True, text=True, timeout=120)
    if lib.rt_device_count() == 0:
        assert r.returncode == 2 and "no CPU fallback" in r.stderr
    else:
        assert r.returncode == 0, r.stderr
        assert (tmp_path / "c.ppm").read_text().startswith("P3\n48 48\n255\n")
