"""CPU tests: the plain-C oracle (oracle/oracle.c) against golden vectors produced by the reference itself
(tests/golden/ref_vectors.bin <- oracle/ref_dump.cpp linked against the unmodified reference).  Bit-exact."""
import numpy as np
import pytest

import oracle_py as O


@pytest.fixture(scope="module")
def G():
    return O.golden()


@pytest.fixture(scope="module")
def T():
    return O.tables()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_hash_and_mixbits(G):
    h = G["hash"].reshape(-1, 13)
    L = O.lib()
    for row in h:
        a, c = int(row[0]), int(row[1])
        buf = np.array([a, c], np.int32)
        assert L.orc_hash_bytes(buf.ctypes.data, 8) == int(row[2]) | (int(row[3]) << 32)
        assert L.orc_mix_bits(((a << 20) ^ c) & 0xffffffffffffffff) == int(row[4]) | (int(row[5]) << 32)
        f3 = row[6:9].copy().view(np.float32)
        assert L.orc_hash_bytes(f3.ctypes.data, 12) == int(row[9]) | (int(row[10]) << 32)
        assert L.orc_hash_bytes(f3[:1].copy().ctypes.data, 4) == int(row[11]) | (int(row[12]) << 32)


@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
def test_samplers(G, T, cfg):
    spp, rx, ry, seed, rs = [int(v) for v in G[f"sampler{cfg}_params"]]
    inp = G[f"sampler{cfg}_in"].reshape(-1, 4)
    z = G[f"sampler{cfg}_zsobol"].reshape(-1, 3)
    ind = G[f"sampler{cfg}_indep"].reshape(-1, 3)
    sob = T["sobol32_2x52"]
    for (px, py, si, dim), zr, ir in zip(inp, z, ind):
        assert (bits(O.zsobol(sob, spp, (rx, ry), seed, rs, px, py, si, dim)) == bits(zr)).all()
        assert (bits(O.independent(seed, px, py, si, dim)) == bits(ir)).all()


def test_intersect_triangle(G):
    i = G["tri_in"].reshape(-1, 16)
    o = G["tri_out"].reshape(-1, 5)
    nhit = 0
    for a, b in zip(i, o):
        hit, out = O.intersect_triangle(a[0:3], a[3:6], a[6], a[7:10], a[10:13], a[13:16])
        assert hit == (b[0] == 1)
        if hit:
            nhit += 1
            assert (bits(out) == bits(b[1:5])).all()
    assert nhit > 500


def test_triangle_badcase_must_miss():
    # the reference's own known-answer test, Triangle.BadCases (shapes_test.cpp:435-450)
    p0, p1, p2 = (-1113.45459, -79.049614, -56.2431908), (-1113.45459, -87.0922699, -56.2431908), (-1113.45459, -79.2090149, -56.2431908)
    o, d = (-1099.77502, -79.1386261, -56.2431908), (-0.999959469, 0.0090017123, 0)
    hit, _ = O.intersect_triangle(o, d, np.inf, p0, p1, p2)
    assert not hit


def test_sampling_routines(G):
    i = G["sampling_in"].reshape(-1, 18)
    o = G["sampling_out"].reshape(-1, 17)
    L = O.lib()
    import ctypes as C
    for a, b in zip(i, o):
        out = np.zeros(4, np.float32)
        L.orc_sample_uniform_disk_concentric(C.c_float(a[0]), C.c_float(a[1]), out.ctypes.data)
        assert (bits(out[:2]) == bits(b[0:2])).all()
        L.orc_sample_cosine_hemisphere(C.c_float(a[0]), C.c_float(a[1]), out.ctypes.data)
        assert (bits(out[:3]) == bits(b[2:5])).all()
        lam = L.orc_sample_visible_wavelengths(C.c_float(a[0]))
        assert bits(np.float32(lam)) == bits(b[5]) and bits(np.float32(L.orc_visible_wavelengths_pdf(C.c_float(lam)))) == bits(b[6])
        v = np.ascontiguousarray(a[2:11]); p = np.ascontiguousarray(a[11:14])
        L.orc_sample_spherical_triangle(v.ctypes.data, p.ctypes.data, C.c_float(a[0]), C.c_float(a[1]), out.ctypes.data)
        assert (bits(out[:3]) == bits(b[7:10])).all() and bits(out[3]) == bits(b[10])
        w = np.ascontiguousarray(a[14:18]); o3 = np.zeros(3, np.float32)
        L.orc_sample_bilinear(C.c_float(a[0]), C.c_float(a[1]), w.ctypes.data, o3.ctypes.data)
        assert (bits(o3) == bits(b[11:14])).all()
        assert bits(np.float32(L.orc_spherical_triangle_area(v.ctypes.data, p.ctypes.data))) == bits(b[14])


def test_spectrum_and_fastexp(G, T):
    i = G["spectrum_in"].reshape(-1, 5)
    o = G["spectrum_out"].reshape(-1, 7)
    L = O.lib()
    import ctypes as C
    scale, coeffs = T["rgb2spec_srgb_scale"], T["rgb2spec_srgb_coeffs"]
    for a, b in zip(i, o):
        assert bits(np.float32(L.orc_fast_exp(C.c_float(a[0])))) == bits(b[0])
        c = np.zeros(3, np.float32)
        rgb = np.ascontiguousarray(a[1:4])
        L.orc_rgb_to_coeffs(scale.ctypes.data, coeffs.ctypes.data, rgb.ctypes.data, c.ctypes.data)
        assert (bits(c) == bits(b[1:4])).all()
        assert bits(np.float32(L.orc_sigmoid_poly(c.ctypes.data, C.c_float(a[4])))) == bits(b[4])   # RGBAlbedoSpectrum(lambda)


def test_bruteforce_closest_hit_equals_reference_bvh(G):
    verts = G["bvh_verts"].reshape(-1, 9)
    rays = G["bvh_rays"].reshape(-1, 8)[:3000]
    hits = G["bvh_hits"].reshape(-1, 12)[:3000]
    prim, bary, t = O.closest_hits(verts, rays)
    occ = O.any_hits(verts, rays)
    assert (prim == hits[:, 0].astype(np.int32)).all()
    m = prim >= 0
    assert (bits(t[m]) == bits(hits[m, 4])).all()
    assert (bits(bary[m]) == bits(hits[m, 1:4])).all()
    assert (occ == hits[:, 5].astype(np.int32)).all()
    assert m.sum() > 500
