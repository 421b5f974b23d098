"""AddressSanitizer + UBSan over the per-thread device functions (host test build) and the host scene code (builders, JSON
writer + reader), on every built-in scene.  compute-sanitizer is closed on the GPU pool; this is the memory-safety check of the code
the kernels are made of (scene flattening, LBVH build bodies, traversal, shading).  CPU only."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_functions_under_asan_ubsan(tmp_path):
    host = os.path.join(REPO, "real-time-ray-tracing-engine_b200", "host")
    exe = str(tmp_path / "asan_emu")
    obj = str(tmp_path / "oracle.o")
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-ffp-contract=off"]
    subprocess.check_call(["gcc", "-c", *flags, os.path.join(REPO, "oracle", "rt_oracle.c"), "-o", obj])
    subprocess.check_call(["g++", "-std=c++17", *flags, "-Wno-unknown-pragmas", os.path.join(REPO, "tests", "emu", "asan_main.cpp"),
                           *[os.path.join(host, f) for f in ("scene_builder.cpp", "scenes.cpp", "host_api.cpp",
                                                             "scene_json.cpp", "cli.cpp")], obj, "-lm", "-o", exe])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", LD_PRELOAD="")
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr
    lines = [l for l in r.stdout.splitlines() if "bvh check" in l]
    assert len(lines) == 6 and all("bvh check 0" in l for l in lines)
