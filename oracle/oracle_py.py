"""ctypes access to the plain-C oracle (oracle/oracle.c) plus readers for the golden-vector / table blobs.
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by
the product package."""
import ctypes as C
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_lib = None


def lib():
    global _lib
    if _lib is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("build_oracle", os.path.join(HERE, "build_oracle.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _lib = C.CDLL(m.build())
        _lib.orc_hash_bytes.restype = C.c_uint64
        _lib.orc_hash_bytes.argtypes = [C.c_void_p, C.c_int]
        _lib.orc_mix_bits.restype = C.c_uint64
        _lib.orc_mix_bits.argtypes = [C.c_uint64]
        _lib.orc_intersect_triangle.restype = C.c_int
        for f in ("orc_sample_visible_wavelengths", "orc_visible_wavelengths_pdf", "orc_fast_exp"):
            getattr(_lib, f).restype = C.c_float
            getattr(_lib, f).argtypes = [C.c_float]
        _lib.orc_spherical_triangle_area.restype = C.c_float
        _lib.orc_sigmoid_poly.restype = C.c_float
        _lib.orc_sigmoid_poly.argtypes = [C.c_void_p, C.c_float]
        V, F, I = C.c_void_p, C.c_float, C.c_int
        _lib.orc_zsobol.argtypes = [V, I, I, I, I, I, I, I, I, I, V]
        _lib.orc_independent.argtypes = [I, I, I, I, I, V]
        _lib.orc_intersect_triangle.argtypes = [V, V, F, V, V, V, V]
        _lib.orc_closest_hits.argtypes = [V, I, V, I, V, V, V]
        _lib.orc_any_hits.argtypes = [V, I, V, I, V]
        _lib.orc_sample_uniform_disk_concentric.argtypes = [F, F, V]
        _lib.orc_sample_cosine_hemisphere.argtypes = [F, F, V]
        _lib.orc_sample_bilinear.argtypes = [F, F, V, V]
        _lib.orc_spherical_triangle_area.argtypes = [V, V]
        _lib.orc_sample_spherical_triangle.argtypes = [V, V, F, F, V]
        _lib.orc_rgb_to_coeffs.argtypes = [V, V, V, V]
    return _lib


def read_blob(path):
    """Reads the tagged blob format written by oracle/ref_dump.cpp -> {name: ndarray}."""
    out = {}
    with open(path, "rb") as f:
        if f.read(8)[:7] != b"BPTTBL1":
            raise ValueError("bad blob")
        while True:
            nm = f.read(32)
            if len(nm) < 32:
                break
            dtype, = struct.unpack("<I", f.read(4))
            count, = struct.unpack("<Q", f.read(8))
            dt = {0: np.float32, 1: np.uint32, 2: np.int32}[dtype]
            out[nm.split(b"\0")[0].decode()] = np.frombuffer(f.read(4 * count), dtype=dt).copy()
    return out


def tables():
    return read_blob(os.path.join(ROOT, "pbrt-v4_b200", "data", "pbrt_tables.bin"))


def golden():
    return read_blob(os.path.join(ROOT, "tests", "golden", "ref_vectors.bin"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def zsobol(sobol, spp, res, seed, randomize, px, py, index, dim):
    out = np.zeros(3, np.float32)
    lib().orc_zsobol(_p(sobol), int(spp), int(res[0]), int(res[1]), int(seed), int(randomize), int(px), int(py), int(index), int(dim), _p(out))
    return out


def independent(seed, px, py, index, dim):
    out = np.zeros(3, np.float32)
    lib().orc_independent(int(seed), int(px), int(py), int(index), int(dim), _p(out))
    return out


def intersect_triangle(o, d, tmax, p0, p1, p2):
    o, d, p0, p1, p2 = [np.ascontiguousarray(v, np.float32) for v in (o, d, p0, p1, p2)]
    out = np.zeros(4, np.float32)
    hit = lib().orc_intersect_triangle(_p(o), _p(d), C.c_float(float(tmax)), _p(p0), _p(p1), _p(p2), _p(out))
    return bool(hit), out


def closest_hits(verts, rays):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    n = rays.shape[0]
    prim = np.zeros(n, np.int32); bary = np.zeros((n, 3), np.float32); t = np.zeros(n, np.float32)
    lib().orc_closest_hits(_p(verts), verts.shape[0], _p(rays), n, _p(prim), _p(bary), _p(t))
    return prim, bary, t


def any_hits(verts, rays):
    verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
    rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
    occ = np.zeros(rays.shape[0], np.int32)
    lib().orc_any_hits(_p(verts), verts.shape[0], _p(rays), rays.shape[0], _p(occ))
    return occ


def cornell_render_space_triangles():
    """The 12 Cornell-box triangles of pbrt-v4_b200/scenes.cornell() in render space (camera-world: world translated by
    the camera position (0, 1, 3.4)), in scene order."""
    quads = [[-0.25, 1.98, -0.25, 0.25, 1.98, -0.25, 0.25, 1.98, 0.25, -0.25, 1.98, 0.25],
             [-1, 0, -1, -1, 0, 1, 1, 0, 1, 1, 0, -1], [-1, 2, -1, 1, 2, -1, 1, 2, 1, -1, 2, 1],
             [-1, 0, -1, 1, 0, -1, 1, 2, -1, -1, 2, -1], [-1, 0, -1, -1, 2, -1, -1, 2, 1, -1, 0, 1],
             [1, 0, -1, 1, 0, 1, 1, 2, 1, 1, 2, -1]]
    cam = np.array([0, 1, 3.4], np.float32)
    tris = []
    for q in quads:
        P = np.array(q, np.float32).reshape(4, 3)
        # renderFromObject(p) = p + (-cam): evaluated as m00*x + m01*y + m02*z + m03 in float (identity rows)
        P = np.stack([(P[:, 0] * np.float32(1) + P[:, 1] * np.float32(0) + P[:, 2] * np.float32(0)) + (-cam[0]),
                      (P[:, 0] * np.float32(0) + P[:, 1] * np.float32(1) + P[:, 2] * np.float32(0)) + (-cam[1]),
                      (P[:, 0] * np.float32(0) + P[:, 1] * np.float32(0) + P[:, 2] * np.float32(1)) + (-cam[2])], axis=1).astype(np.float32)
        for idx in ((0, 1, 2), (0, 2, 3)):
            tris.append(np.concatenate([P[idx[0]], P[idx[1]], P[idx[2]]]))
    return np.array(tris, np.float32)
