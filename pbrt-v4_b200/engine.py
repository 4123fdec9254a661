"""ctypes binding of the C ABI (include/b200pt.h) and a thin host-side mirror of the reference's
WavefrontPathIntegrator driver (wavefront/wavefront.cpp:14-70 RenderWavefront: construct, Render(), write image).

The CUDA library is REQUIRED: importing this module builds nothing and falls back to nothing -- `Engine()` raises
if libb200pt.so is missing or no sm_100 device is present.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BPT_LIB") or os.path.join(_HERE, "libb200pt.so")   # BPT_LIB: tuning-variant builds (scripts/sweep.py)
_lib = None


class BptStats(C.Structure):
    _fields_ = [("seconds_build", C.c_double), ("seconds_render", C.c_double),
                ("camera_rays", C.c_uint64), ("indirect_rays", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("counted_rays", C.c_uint64),
                ("nodes_visited_shadow", C.c_uint64), ("tris_tested_shadow", C.c_uint64), ("counted_rays_shadow", C.c_uint64),
                ("bvh_nodes", C.c_int), ("bvh_triangles", C.c_int), ("bvh_depth", C.c_int),
                ("bytes_bvh_nodes", C.c_uint64), ("bytes_bvh_tris", C.c_uint64), ("kernel_launches", C.c_int),
                ("ms_trace_closest", C.c_double), ("ms_trace_shadow", C.c_double), ("ms_shade", C.c_double),
                ("ms_other", C.c_double), ("launches_trace_closest", C.c_int), ("launches_trace_shadow", C.c_int),
                ("ms_medium", C.c_double)]


EXPORTS = """bpt_create bpt_destroy bpt_last_error bpt_scene_begin bpt_scene_set_camera bpt_scene_set_film
bpt_scene_set_filter bpt_scene_set_sampler bpt_scene_set_integrator bpt_scene_set_materials bpt_scene_set_area_lights
bpt_scene_add_mesh bpt_scene_add_quadric bpt_scene_add_bilinear_mesh bpt_scene_add_instance bpt_scene_add_light bpt_scene_add_medium bpt_scene_set_camera_medium
bpt_scene_commit bpt_load_pbrt bpt_load_pbrt_ex bpt_frontend_error bpt_render bpt_render_bands bpt_film_resolve_gbuffer bpt_film_clear bpt_stats_reset bpt_film_read bpt_film_write
bpt_film_copy_to_device bpt_film_copy_from_device bpt_film_resolve bpt_film_dims bpt_get_stats bpt_trace_closest
bpt_trace_any bpt_trace_closest_device bpt_probe_sampler bpt_probe_intersect_triangle bpt_render_stream bpt_render_bands_stream
bpt_sync bpt_sampler_spp bpt_scene_set_transform_space bpt_set_option bpt_film_write_exr bpt_write_exr_file""".split()


def load_library():
    """dlopen libb200pt.so and declare prototypes.  Raises (never falls back) if the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python pbrt-v4_b200/build.py` "
                           "(there is no CPU or Python fallback for the render path)")
    lib = C.CDLL(LIB_PATH)
    lib.bpt_create.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.bpt_destroy.argtypes = [C.c_void_p]
    lib.bpt_destroy.restype = None
    lib.bpt_last_error.argtypes = [C.c_void_p]
    lib.bpt_last_error.restype = C.c_char_p
    lib.bpt_frontend_error.restype = C.c_char_p
    lib.bpt_load_pbrt_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    lib.bpt_render.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.bpt_render_bands.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.bpt_render_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.c_void_p]
    lib.bpt_render_bands_stream.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    lib.bpt_sync.argtypes = [C.c_void_p]
    lib.bpt_film_write_exr.argtypes = [C.c_void_p, C.c_char_p]
    lib.bpt_write_exr_file.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.c_void_p]
    lib.bpt_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_longlong]
    lib.bpt_sampler_spp.argtypes = [C.c_void_p]
    lib.bpt_film_clear.argtypes = [C.c_void_p]
    lib.bpt_stats_reset.argtypes = [C.c_void_p]
    lib.bpt_film_read.argtypes = [C.c_void_p, C.c_void_p]
    lib.bpt_film_write.argtypes = [C.c_void_p, C.c_void_p]
    lib.bpt_film_copy_to_device.argtypes = [C.c_void_p, C.c_void_p]
    lib.bpt_film_copy_from_device.argtypes = [C.c_void_p, C.c_void_p]
    lib.bpt_film_resolve.argtypes = [C.c_void_p, C.c_void_p]
    lib.bpt_film_resolve_gbuffer.argtypes = [C.c_void_p, C.c_void_p]
    lib.bpt_film_dims.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.bpt_get_stats.argtypes = [C.c_void_p, C.POINTER(BptStats)]
    lib.bpt_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bpt_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.bpt_trace_closest_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.bpt_probe_sampler.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.bpt_probe_intersect_triangle.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    _lib = lib
    return lib


class BptError(RuntimeError):
    pass


class Engine:
    """One engine per GPU (mirrors one WavefrontPathIntegrator instance)."""

    def __init__(self, device=0, tables_path=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.bpt_create(device, tables_path.encode() if tables_path else None, C.byref(h))
        if rc != 0:
            raise BptError(f"bpt_create failed ({rc}): {self.lib.bpt_last_error(None).decode()}")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.bpt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            msg = self.lib.bpt_last_error(self.h).decode()
            raise BptError(f"{what} failed ({rc}): {msg}")

    # ---- scene
    def load_pbrt(self, path, spp=0, seed=0):
        rc = self.lib.bpt_load_pbrt_ex(self.h, os.fspath(path).encode(), int(spp), int(seed))
        if rc != 0:
            raise BptError(f"load_pbrt({path}) failed: {self.lib.bpt_frontend_error().decode()} / "
                           f"{self.lib.bpt_last_error(self.h).decode()}")

    # ---- render
    def render(self, first_sample, n_samples, pixel_bounds=None, count_stats=False, profile=False):
        pb = None
        if pixel_bounds is not None:
            pb = (C.c_int * 4)(*[int(v) for v in pixel_bounds])
        self._ck(self.lib.bpt_render(self.h, int(first_sample), int(n_samples), pb, (1 if count_stats else 0) | (2 if profile else 0)), "bpt_render")

    def render_async(self, first_sample, n_samples, pixel_bounds=None, stream=0):
        """Enqueue a render on `stream` (a cudaStream_t handle as int; 0 = the engine's stream) and return; sync() waits."""
        pb = (C.c_int * 4)(*[int(v) for v in pixel_bounds]) if pixel_bounds is not None else None
        self._ck(self.lib.bpt_render_stream(self.h, int(first_sample), int(n_samples), pb, 4, C.c_void_p(stream)), "bpt_render_stream")

    def set_option(self, name, value):
        self._ck(self.lib.bpt_set_option(self.h, name.encode(), int(value)), "bpt_set_option")

    def write_exr(self, path):
        """Film "rgb" -> R G B, Film "gbuffer" -> the 25 reference channels, as an uncompressed FLOAT OpenEXR file."""
        self._ck(self.lib.bpt_film_write_exr(self.h, os.fspath(path).encode()), "bpt_film_write_exr")

    def sync(self):
        self._ck(self.lib.bpt_sync(self.h), "bpt_sync")

    def sampler_spp(self):
        v = self.lib.bpt_sampler_spp(self.h)
        if v < 0:
            raise BptError("bpt_sampler_spp: scene not committed")
        return v

    def film_clear(self):
        self._ck(self.lib.bpt_film_clear(self.h), "bpt_film_clear")

    def stats_reset(self):
        self._ck(self.lib.bpt_stats_reset(self.h), "bpt_stats_reset")

    def dims(self):
        x, y = C.c_int(), C.c_int()
        self._ck(self.lib.bpt_film_dims(self.h, C.byref(x), C.byref(y)), "bpt_film_dims")
        return x.value, y.value

    def film(self):
        """Film accumulators as float64 array [yres, xres, 4] = rgbSum[3], weightSum."""
        x, y = self.dims()
        out = np.empty((y, x, 4), dtype=np.float64)
        self._ck(self.lib.bpt_film_read(self.h, out.ctypes.data), "bpt_film_read")
        return out

    def film_write(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        self._ck(self.lib.bpt_film_write(self.h, arr.ctypes.data), "bpt_film_write")

    def film_to_device_ptr(self, ptr):
        self._ck(self.lib.bpt_film_copy_to_device(self.h, C.c_void_p(ptr)), "bpt_film_copy_to_device")

    def film_from_device_ptr(self, ptr):
        self._ck(self.lib.bpt_film_copy_from_device(self.h, C.c_void_p(ptr)), "bpt_film_copy_from_device")

    def render_bands(self, first_sample, n_samples, band_rows, n_parts, part, count_stats=False, profile=False):
        """Render the interleaved scanline bands of part `part` of `n_parts` in one call (multi-GPU tiling)."""
        self._ck(self.lib.bpt_render_bands(self.h, int(first_sample), int(n_samples), int(band_rows), int(n_parts), int(part),
                                           (1 if count_stats else 0) | (2 if profile else 0)), "bpt_render_bands")

    def image(self):
        """Final RGB image [yres, xres, 3] float32 (RGBFilm::GetPixelRGB)."""
        x, y = self.dims()
        out = np.empty((y, x, 3), dtype=np.float32)
        self._ck(self.lib.bpt_film_resolve(self.h, out.ctypes.data), "bpt_film_resolve")
        return out

    GBUFFER_CHANNELS = ["R", "G", "B", "Albedo.R", "Albedo.G", "Albedo.B", "P.X", "P.Y", "P.Z", "dzdx", "dzdy", "N.X", "N.Y", "N.Z",
                        "Ns.X", "Ns.Y", "Ns.Z", "u", "v", "Variance.R", "Variance.G", "Variance.B",
                        "RelativeVariance.R", "RelativeVariance.G", "RelativeVariance.B"]

    def gbuffer(self):
        """GBufferFilm image [yres, xres, 25] float32 in the reference's channel order (GBufferFilm::GetImage)."""
        x, y = self.dims()
        out = np.empty((y, x, 25), dtype=np.float32)
        self._ck(self.lib.bpt_film_resolve_gbuffer(self.h, out.ctypes.data), "bpt_film_resolve_gbuffer")
        return out

    def stats(self):
        s = BptStats()
        self._ck(self.lib.bpt_get_stats(self.h, C.byref(s)), "bpt_get_stats")
        return {k: getattr(s, k) for k, _ in BptStats._fields_}

    # ---- ray-level
    def trace_closest(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        hits = np.empty((n, 4), dtype=np.float32)
        t = np.empty(n, dtype=np.float32)
        self._ck(self.lib.bpt_trace_closest(self.h, rays.ctypes.data, n, hits.ctypes.data, t.ctypes.data), "bpt_trace_closest")
        return hits[:, 0].copy().view(np.int32), hits[:, 1:4].copy(), t

    def trace_any(self, rays):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        occ = np.empty(n, dtype=np.int32)
        self._ck(self.lib.bpt_trace_any(self.h, rays.ctypes.data, n, occ.ctypes.data), "bpt_trace_any")
        return occ

    def probe_sampler(self, pixels, sample_index, dim):
        pixels = np.ascontiguousarray(pixels, dtype=np.int32).reshape(-1, 2)
        n = pixels.shape[0]
        si = np.ascontiguousarray(sample_index, dtype=np.int32)
        dm = np.ascontiguousarray(dim, dtype=np.int32)
        o1 = np.empty(n, dtype=np.float32)
        o2 = np.empty((n, 2), dtype=np.float32)
        self._ck(self.lib.bpt_probe_sampler(self.h, pixels.ctypes.data, si.ctypes.data, dm.ctypes.data, n, o1.ctypes.data,
                                            o2.ctypes.data), "bpt_probe_sampler")
        return o1, o2

    def probe_intersect_triangle(self, rays7, tris9):
        rays7 = np.ascontiguousarray(rays7, dtype=np.float32).reshape(-1, 7)
        tris9 = np.ascontiguousarray(tris9, dtype=np.float32).reshape(-1, 9)
        n = rays7.shape[0]
        out = np.empty((n, 5), dtype=np.float32)
        self._ck(self.lib.bpt_probe_intersect_triangle(self.h, rays7.ctypes.data, tris9.ctypes.data, n, out.ctypes.data),
                 "bpt_probe_intersect_triangle")
        return out


def render_pbrt(path, spp=0, seed=0, device=0):
    """RenderWavefront equivalent (wavefront/wavefront.cpp:14-70): load the scene, render all of its samples (`spp` > 0 overrides
    the file's pixelsamples like --spp), return (image, stats)."""
    e = Engine(device)
    try:
        e.load_pbrt(path, spp=spp, seed=seed)
        e.render(0, e.sampler_spp(), None)
        return e.image(), e.stats()
    finally:
        e.close()
