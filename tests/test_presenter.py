"""The dynamic camera's SDL3 window (host/presenter.cpp) against a test double of libSDL3 (tests/emu/fake_sdl3.c):
the calls DynamicCamera makes (core/camera/DynamicCamera.cpp:62-91,196-306), its key handling, and the failure
paths when SDL3 is absent.  The product opens SDL3 with dlopen, so none of this needs SDL3 or a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from rt_b200 import host

HERE = os.path.dirname(os.path.abspath(__file__))
FAKE = os.path.join(HERE, "emu", "libfake_sdl3.so")


class rth_input(C.Structure):
    _fields_ = [("quit", C.c_int), ("spp_delta", C.c_int), ("move_x", C.c_int), ("move_z", C.c_int), ("moved", C.c_int)]


@pytest.fixture()
def lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "emu"), "libfake_sdl3.so"])
    h = host.load_host_library()
    h.rth_presenter_open.restype = C.c_void_p
    h.rth_presenter_open.argtypes = [C.c_int, C.c_int, C.c_char_p]
    h.rth_presenter_poll.argtypes = [C.c_void_p, C.POINTER(rth_input)]
    h.rth_presenter_present.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
    h.rth_presenter_close.argtypes = [C.c_void_p]
    h.rth_last_error.restype = C.c_char_p
    return h


def test_window_calls_and_keys(lib, tmp_path, monkeypatch):
    log, frame = tmp_path / "sdl.log", tmp_path / "frame.bin"
    monkeypatch.setenv("RT_SDL3_LIB", FAKE)
    monkeypatch.setenv("FAKE_SDL_LOG", str(log))
    monkeypatch.setenv("FAKE_SDL_FRAME", str(frame))
    monkeypatch.setenv("FAKE_SDL_KEYS", ".wd=-ase")
    W, H = 48, 20
    p = lib.rth_presenter_open(W, H, b"Dynamic Camera")
    assert p, lib.rth_last_error()
    seen = []
    for _ in range(9):
        i = rth_input()
        assert lib.rth_presenter_poll(p, C.byref(i)) == 0
        seen.append((i.quit, i.spp_delta, i.move_x, i.move_z, i.moved))
    assert seen == [(0, 0, 0, 0, 0),   # .
                    (0, 0, 0, 1, 1),   # w held: +z
                    (0, 0, 1, 0, 1),   # d held: +x
                    (0, 1, 0, 0, 0),   # = pressed: one more sample per pixel
                    (0, -1, 0, 0, 0),  # - pressed
                    (0, 0, -1, 0, 1),  # a held: -x
                    (0, 0, 0, -1, 1),  # s held: -z
                    (1, 0, 0, 0, 0),   # ESC
                    (1, 0, 0, 0, 0)]   # script over: window closed
    rgb = np.random.default_rng(3).integers(0, 256, W * H * 3, dtype=np.uint8)
    assert lib.rth_presenter_present(p, rgb.ctypes.data_as(C.c_void_p), b"12.5 fps") == 0
    lib.rth_presenter_close(p)
    assert np.array_equal(np.fromfile(frame, dtype=np.uint8), rgb)  # the whole frame, pitch = 3 * width
    calls = log.read_text().splitlines()
    assert calls[0] == "SDL_Init 0x20"  # SDL_INIT_VIDEO
    assert calls[1] == 'SDL_CreateWindow "Dynamic Camera" 48 20 0'
    assert calls[2] == "SDL_CreateRenderer 1 (null)"
    assert calls[3] == "SDL_CreateTexture 1 0x17101803 1 48 20"  # RGB24, streaming
    assert calls[4:9] == ["SDL_UpdateTexture 1 rect=0 pitch=144", 'SDL_SetWindowTitle 1 "12.5 fps"', "SDL_RenderClear 1",
                          "SDL_RenderTexture 1 1 0 0", "SDL_RenderPresent 1"]
    assert calls[9:] == ["SDL_DestroyTexture 1", "SDL_DestroyRenderer 1", "SDL_DestroyWindow 1", "SDL_Quit"]


def test_missing_sdl3_is_reported(lib, tmp_path, monkeypatch):
    monkeypatch.setenv("RT_SDL3_LIB", str(tmp_path / "no-such-libSDL3.so"))
    assert not lib.rth_presenter_open(32, 32, b"x")
    assert b"SDL3 is not available" in lib.rth_last_error() and b"--frames" in lib.rth_last_error()
    # a library that is not SDL3: every missing entry point is named
    monkeypatch.setenv("RT_SDL3_LIB", os.path.join(HERE, "emu", "libemu.so"))
    if os.path.exists(os.path.join(HERE, "emu", "libemu.so")):
        assert not lib.rth_presenter_open(32, 32, b"x")
        assert b"SDL_CreateWindow" in lib.rth_last_error()
    monkeypatch.setenv("RT_SDL3_LIB", FAKE)
    monkeypatch.setenv("FAKE_SDL_FAIL_INIT", "1")
    assert not lib.rth_presenter_open(32, 32, b"x")
    assert b"SDL_Init failed" in lib.rth_last_error()
    assert not lib.rth_presenter_open(0, 32, b"x")
