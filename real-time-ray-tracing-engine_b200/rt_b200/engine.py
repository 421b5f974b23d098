"""Thin object wrappers over the C ABI of librt_b200.so (include/rt_b200.h).

The class names follow the reference's render drivers: a `Context` is one GPU, a `Scene` the uploaded
world + BVH (initialize_cuda_scene), a `Film` the accumulation buffer of a Static/Dynamic camera, and
`render_static` / `render_accumulate` the two launch wrappers the reference's cameras call
(core/camera/CameraKernelWrappers.cuh:15-32).  All compute happens in the CUDA library; importing this
module without it raises (no CPU fallback).
"""
import ctypes as C

import numpy as np

from . import abi


class Context:
    def __init__(self, device=0):
        self.lib = abi.load_library()
        h = C.c_void_p()
        abi.check(self.lib, self.lib.rt_context_create(device, C.byref(h)), "rt_context_create")
        self._h = h
        self.device = device

    @property
    def stream(self):
        """cudaStream_t of the context as an integer."""
        return self.lib.rt_context_stream(self._h)

    def synchronize(self):
        abi.check(self.lib, self.lib.rt_context_synchronize(self._h), "rt_context_synchronize")

    def counters(self):
        c = abi.rt_counters()
        abi.check(self.lib, self.lib.rt_get_counters(self._h, C.byref(c)), "rt_get_counters")
        return c

    def set_stage_timing(self, enable):
        abi.check(self.lib, self.lib.rt_context_set_stage_timing(self._h, int(enable)), "rt_context_set_stage_timing")

    def stage_times(self):
        """(ms[5], launches[5]) for generate / extend / shade / accumulate / tail since the last call."""
        ms = (C.c_double * 5)()
        n = (C.c_uint64 * 5)()
        abi.check(self.lib, self.lib.rt_get_stage_times(self._h, ms, n), "rt_get_stage_times")
        return list(ms), list(n)

    def set_graph(self, enable):
        """Render passes as one CUDA-graph launch each (captured per pass, executable graph updated in place)."""
        abi.check(self.lib, self.lib.rt_context_set_graph(self._h, int(enable)), "rt_context_set_graph")

    def set_stats(self, enable):
        """Instrumented extend / tail kernels: counters() then reports node visits and primitive tests."""
        abi.check(self.lib, self.lib.rt_context_set_stats(self._h, int(enable)), "rt_context_set_stats")

    def queue_lengths(self, n):
        out = (C.c_uint32 * n)()
        abi.check(self.lib, self.lib.rt_get_queue_lengths(self._h, out, n), "rt_get_queue_lengths")
        return list(out)

    def set_audit(self, enable):
        """Parity audit: every segment of the following render passes is also traced by the FP64 parity traversal."""
        abi.check(self.lib, self.lib.rt_context_set_audit(self._h, int(enable)), "rt_context_set_audit")

    def audit(self):
        a = abi.rt_audit()
        abi.check(self.lib, self.lib.rt_get_audit(self._h, C.byref(a)), "rt_get_audit")
        return a

    def audit_samples(self, max_samples=4096):
        buf = (abi.rt_audit_sample * max_samples)()
        n = self.lib.rt_get_audit_samples(self._h, buf, max_samples)
        if n < 0:
            raise abi.RtError("rt_get_audit_samples failed")
        return buf[:n]

    def reset_counters(self):
        abi.check(self.lib, self.lib.rt_reset_counters(self._h), "rt_reset_counters")

    def close(self):
        if self._h:
            self.lib.rt_context_destroy(self._h)
            self._h = None


class Scene:
    def __init__(self, ctx, desc):
        """desc: rt_scene_desc or a pointer to one (e.g. HostScene.desc)."""
        self.ctx = ctx
        self.lib = ctx.lib
        h = C.c_void_p()
        ptr = desc if isinstance(desc, C.POINTER(abi.rt_scene_desc)) else C.pointer(desc)
        abi.check(self.lib, self.lib.rt_scene_create(ctx._h, ptr, C.byref(h)), "rt_scene_create")
        self._h = h

    def info(self):
        i = abi.rt_scene_info()
        abi.check(self.lib, self.lib.rt_scene_get_info(self._h, C.byref(i)), "rt_scene_get_info")
        return i

    def update_spheres(self, first, spheres):
        """Animated scenes: replace spheres [first, first + len(spheres)) and refit the BVH.
        spheres: ctypes array of rt_sphere."""
        abi.check(self.lib, self.lib.rt_scene_update_spheres(self._h, first, len(spheres), spheres), "rt_scene_update_spheres")

    def update_quads(self, first, quads):
        """The same for quads; quads: ctypes array of rt_quad."""
        abi.check(self.lib, self.lib.rt_scene_update_quads(self._h, first, len(quads), quads), "rt_scene_update_quads")

    def trace(self, rays, mode=abi.RT_TRACE_EXACT_F64, seed=0):
        """rays: ctypes array of rt_ray.  Returns a ctypes array of rt_hit."""
        n = len(rays)
        hits = (abi.rt_hit * n)()
        abi.check(self.lib, self.lib.rt_trace_rays(self._h, rays, n, mode, seed, hits), "rt_trace_rays")
        return hits

    def close(self):
        if self._h:
            self.lib.rt_scene_destroy(self._h)
            self._h = None


def camera_from_config(cfg):
    lib = abi.load_library()
    cam = abi.rt_camera()
    abi.check(lib, lib.rt_camera_init(C.byref(cfg), C.byref(cam)), "rt_camera_init")
    return cam


class Film:
    def __init__(self, ctx, width, height, rank=0, n_ranks=1, tile_rows=8, external_accum=None):
        self.ctx = ctx
        self.lib = ctx.lib
        self.width, self.height = width, height
        self.rank, self.n_ranks, self.tile_rows = rank, n_ranks, tile_rows
        h = C.c_void_p()
        ext = C.c_void_p(external_accum) if external_accum else None
        abi.check(self.lib, self.lib.rt_film_create(ctx._h, width, height, rank, n_ranks, tile_rows, ext, C.byref(h)),
                  "rt_film_create")
        self._h = h

    @property
    def owned_pixels(self):
        return self.lib.rt_film_owned_pixels(self._h)

    @property
    def samples(self):
        return self.lib.rt_film_samples(self._h)

    @property
    def device_ptr(self):
        return self.lib.rt_film_device_ptr(self._h)

    def clear(self):
        abi.check(self.lib, self.lib.rt_film_clear(self._h), "rt_film_clear")

    def read_rgb(self, scale):
        out = np.empty((self.owned_pixels, 3), dtype=np.float32)
        abi.check(self.lib, self.lib.rt_film_read_rgb(self._h, scale, out.ctypes.data_as(C.POINTER(C.c_float))),
                  "rt_film_read_rgb")
        return out

    def resolve_rgb8(self, scale, out=None):
        if out is None:
            out = np.empty((self.owned_pixels, 3), dtype=np.uint8)
        abi.check(self.lib, self.lib.rt_film_resolve_rgb8(self._h, scale, out.ctypes.data_as(C.POINTER(C.c_uint8))),
                  "rt_film_resolve_rgb8")
        return out

    def owned_rows(self):
        """Global scanline index of every owned row, in storage order."""
        rows = []
        n_tiles = (self.height + self.tile_rows - 1) // self.tile_rows
        for t in range(self.rank, n_tiles, self.n_ranks):
            rows.extend(range(t * self.tile_rows, min((t + 1) * self.tile_rows, self.height)))
        return np.asarray(rows, dtype=np.int64)

    def close(self):
        if self._h:
            self.lib.rt_film_destroy(self._h)
            self._h = None


def render_accumulate(scene, camera, film, s_i, s_j, sqrt_spp, max_depth, seed):
    """One progressive frame (cuda_dynamic_render_tile_wrapper semantics, whole frame).  Asynchronous."""
    abi.check(scene.lib, scene.lib.rt_render_accumulate(scene._h, C.byref(camera), film._h, s_i, s_j, sqrt_spp,
                                                        max_depth, seed), "rt_render_accumulate")


def render_static(scene, camera, film, sqrt_spp, max_depth, seed):
    """All sqrt_spp^2 strata (cuda_static_render_wrapper semantics).  Asynchronous."""
    abi.check(scene.lib, scene.lib.rt_render_static(scene._h, C.byref(camera), film._h, sqrt_spp, max_depth, seed),
              "rt_render_static")


def render_strata(scene, camera, film, first_stratum, n_strata, sqrt_spp, max_depth, seed):
    """n_strata consecutive strata in one wavefront pass, added to the film.  Asynchronous."""
    abi.check(scene.lib, scene.lib.rt_render_strata(scene._h, C.byref(camera), film._h, first_stratum, n_strata,
                                                    sqrt_spp, max_depth, seed), "rt_render_strata")


class Frame:
    """The displayed RGB8 frame of a (multi-GPU) render in its owner's memory (include/rt_b200.h, "Displayed
    frames"): every rank stores its own tiles straight into it (`present`), the owner waits / downloads."""

    def __init__(self, ctx, width, height, n_ranks=1, ipc_handle=None, attach_to=None):
        self.ctx, self.lib = ctx, ctx.lib
        self.width, self.height, self.n_ranks = width, height, n_ranks
        h = C.c_void_p()
        if ipc_handle is not None:  # another process's frame
            buf = (C.c_ubyte * 64).from_buffer_copy(bytes(ipc_handle))
            abi.check(self.lib, self.lib.rt_frame_open(ctx._h, buf, width, height, n_ranks, C.byref(h)), "rt_frame_open")
        elif attach_to is not None:  # another GPU of this process
            abi.check(self.lib, self.lib.rt_frame_attach(ctx._h, attach_to._h, C.byref(h)), "rt_frame_attach")
        else:
            abi.check(self.lib, self.lib.rt_frame_create(ctx._h, width, height, n_ranks, C.byref(h)), "rt_frame_create")
        self._h = h

    def export(self):
        buf = (C.c_ubyte * 64)()
        abi.check(self.lib, self.lib.rt_frame_export(self._h, buf), "rt_frame_export")
        return bytes(buf)

    @property
    def device_ptr(self):
        return self.lib.rt_frame_device_ptr(self._h)

    def present(self, film, scale):
        abi.check(self.lib, self.lib.rt_film_present(film._h, scale, self._h), "rt_film_present")

    def wait(self):
        abi.check(self.lib, self.lib.rt_frame_wait(self._h), "rt_frame_wait")

    def wait_release(self):
        abi.check(self.lib, self.lib.rt_frame_wait_release(self._h), "rt_frame_wait_release")

    def release(self):
        abi.check(self.lib, self.lib.rt_frame_release(self._h), "rt_frame_release")

    def download(self, host_ptr):
        abi.check(self.lib, self.lib.rt_frame_download(self._h, C.c_void_p(host_ptr)), "rt_frame_download")

    def download_wait(self):
        abi.check(self.lib, self.lib.rt_frame_download_wait(self._h), "rt_frame_download_wait")

    def error(self):
        return self.lib.rt_frame_error(self._h)

    def close(self):
        if self._h:
            self.lib.rt_frame_destroy(self._h)
            self._h = None
