"""The `raytracer` host program (reference CLI in front of the B200 backend) end to end on a GPU: PPM output
format, agreement with the library path, JSON scene input, dynamic (headless) camera, multi-GPU flag."""
import os
import subprocess

import numpy as np
import pytest

from rt_b200 import abi, engine, host

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(REPO, "real-time-ray-tracing-engine_b200", "host", "raytracer")


def read_ppm(path):
    with open(path) as f:
        tokens = f.read().split()
    assert tokens[0] == "P3" and tokens[3] == "255"
    w, h = int(tokens[1]), int(tokens[2])
    return np.array(tokens[4:], dtype=np.int64).reshape(h, w, 3).astype(np.uint8)


def run(args, cwd):
    return subprocess.run([EXE] + args, cwd=cwd, capture_output=True, text=True, timeout=300)


def test_static_render_matches_the_library(tmp_path):
    r = run(["--camera", "static", "--scene", "spheres", "--width", "160", "--samples", "16", "--depth", "10", "--output",
             "a.ppm", "-g", "-b", "-p"], tmp_path)
    assert r.returncode == 0, r.stderr
    img = read_ppm(tmp_path / "output" / "a.ppm")
    assert img.shape == (90, 160, 3)
    # same render through the ctypes path: identical bytes (same scene generator, same Philox keys)
    ctx = engine.Context(0)
    hs = host.HostScene.builtin("spheres", 1234)
    cam = engine.camera_from_config(hs.camera_config(160, 16, 10))
    scene = engine.Scene(ctx, hs.desc)
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    engine.render_static(scene, cam, film, 4, 10, 1234)
    want = film.resolve_rgb8(1.0 / 16).reshape(90, 160, 3)
    assert np.array_equal(img, want)
    film.close()
    scene.close()
    ctx.close()


def test_json_scene_and_dynamic_camera(tmp_path):
    hs = host.HostScene.builtin("cornell", 1234)
    hs.save_json(str(tmp_path / "cornell.json"))
    a = run(["--scene", str(tmp_path / "cornell.json"), "--width", "64", "--samples", "4", "--depth", "6", "--output", "json.ppm"],
            tmp_path)
    b = run(["--scene", "cornell", "--width", "64", "--samples", "4", "--depth", "6", "--output", "builtin.ppm"], tmp_path)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert np.array_equal(read_ppm(tmp_path / "output" / "json.ppm"), read_ppm(tmp_path / "output" / "builtin.ppm"))
    # dynamic camera, headless: one stratum per frame, all strata == the static render of the same spp
    d = run(["--camera", "dynamic", "--scene", "cornell", "--width", "64", "--samples", "4", "--depth", "6", "--output",
             "dyn.ppm"], tmp_path)
    assert d.returncode == 0, d.stderr
    assert "4 progressive frames" in d.stderr
    assert np.array_equal(read_ppm(tmp_path / "output" / "dyn.ppm"), read_ppm(tmp_path / "output" / "builtin.ppm"))


def test_cli_errors(tmp_path):
    assert run(["--bogus"], tmp_path).returncode != 0
    assert run(["--scene", "nope"], tmp_path).returncode != 0
    h = run(["--help"], tmp_path)
    assert h.returncode == 0 and "--camera [static|dynamic]" in h.stdout
    lib = abi.load_library()
    too_many = lib.rt_device_count() + 1
    r = run(["--scene", "cornell", "--width", "32", "--samples", "1", "--gpus", str(too_many)], tmp_path)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr


def test_scripted_camera_move_restarts_the_accumulation(tmp_path):
    """DynamicCamera::handle_events (DynamicCamera.cpp:204-278): a movement key shifts lookfrom / lookat by 10
    units, clears the accumulation and re-initialises the camera.  Frames 0-2 accumulate, frame 3 has 'd' down
    (+10 in x), so the last image is the average of frames 3-5 seen from the moved camera."""
    r = run(["--camera", "dynamic", "--scene", "cornell", "--width", "64", "--samples", "9", "--depth", "5", "--frames", "6",
             "--keys", "...d", "--output", "moved.ppm"], tmp_path)
    assert r.returncode == 0, r.stderr
    assert "1 camera move(s), 3 sample(s)" in r.stderr
    ctx = engine.Context(0)
    hs = host.HostScene.builtin("cornell", 1234)
    cfg = hs.camera_config(64, 9, 5)
    cfg.lookfrom[0] += 10.0
    cfg.lookat[0] += 10.0
    cam = engine.camera_from_config(cfg)
    scene = engine.Scene(ctx, hs.desc)
    film = engine.Film(ctx, cam.image_width, cam.image_height)
    for s in range(3):
        engine.render_accumulate(scene, cam, film, s % 3, s // 3, 3, 5, 1234)
    want = film.resolve_rgb8(1.0 / 3).reshape(cam.image_height, cam.image_width, 3)
    assert np.array_equal(read_ppm(tmp_path / "output" / "moved.ppm"), want)
    film.close()
    scene.close()
    ctx.close()


def test_multi_gpu_flag_gives_the_same_bytes(tmp_path):
    """--gpus N: tiles interleaved over the devices, RGB8 tiles gathered over NVLink peer copies; the frame is
    byte-identical to the single-GPU one (Philox keyed by the global pixel, FP64 to_byte on every device)."""
    lib = abi.load_library()
    n = min(lib.rt_device_count(), 4)
    if n < 2:
        pytest.skip("needs at least two GPUs")
    args = ["--scene", "final", "--width", "192", "--samples", "4", "--depth", "12"]
    a = run(args + ["--output", "one.ppm"], tmp_path)
    b = run(args + ["--gpus", str(n), "--output", "many.ppm"], tmp_path)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert np.array_equal(read_ppm(tmp_path / "output" / "one.ppm"), read_ppm(tmp_path / "output" / "many.ppm"))
    d = run(["--camera", "dynamic", "--gpus", str(n), "--output", "dyn.ppm"] + args, tmp_path)
    assert d.returncode == 0, d.stderr
    assert np.array_equal(read_ppm(tmp_path / "output" / "dyn.ppm"), read_ppm(tmp_path / "output" / "one.ppm"))


def test_window_mode_shows_the_frames_the_headless_mode_writes(tmp_path):
    """Dynamic camera with a window (a test double of libSDL3, tests/emu/fake_sdl3.c): keys come from SDL, every
    frame goes through SDL_UpdateTexture, ESC ends the loop.  Same key script headless -> same last frame."""
    fake = os.path.join(REPO, "tests", "emu", "libfake_sdl3.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(REPO, "tests", "emu"), "libfake_sdl3.so"])
    env = dict(os.environ, RT_SDL3_LIB=fake, FAKE_SDL_LOG=str(tmp_path / "sdl.log"), FAKE_SDL_FRAME=str(tmp_path / "frame.bin"),
               FAKE_SDL_KEYS="..d..e")
    args = ["--camera", "dynamic", "--scene", "cornell", "--width", "64", "--samples", "16", "--depth", "5"]
    w = subprocess.run([EXE] + args + ["--output", "win.ppm"], cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env)
    assert w.returncode == 0, w.stderr
    assert "5 progressive frames" in w.stderr and "1 camera move(s), 3 sample(s)" in w.stderr
    h = run(args + ["--frames", "5", "--keys", "..d..", "--output", "headless.ppm"], tmp_path)
    assert h.returncode == 0, h.stderr
    want = read_ppm(tmp_path / "output" / "headless.ppm")
    assert np.array_equal(read_ppm(tmp_path / "output" / "win.ppm"), want)
    shown = np.fromfile(tmp_path / "frame.bin", dtype=np.uint8).reshape(want.shape)
    assert np.array_equal(shown, want)  # what SDL_UpdateTexture received last
    calls = (tmp_path / "sdl.log").read_text().splitlines()
    assert sum(c.startswith("SDL_UpdateTexture") for c in calls) == 5 and calls[-1] == "SDL_Quit"
    assert any("3/16 samples" in c for c in calls if c.startswith("SDL_SetWindowTitle"))
    # without SDL3 the same command falls back to the headless loop and says so
    env2 = dict(os.environ, RT_SDL3_LIB=str(tmp_path / "missing.so"))
    f = subprocess.run([EXE] + ["--camera", "dynamic", "--scene", "cornell", "--width", "32", "--samples", "4", "--depth", "3"],
                       cwd=tmp_path, capture_output=True, text=True, timeout=300, env=env2)
    assert f.returncode == 0 and "rendering headless" in f.stderr and "4 progressive frames" in f.stderr
