"""PFM read/write (the reference's Image::WritePFM layout, util/image.cpp:1551-1815: bottom-to-top rows,
little-endian when the scale is negative) and image-difference metrics used by the parity tests."""
import numpy as np


def read_pfm(path):
    with open(path, "rb") as f:
        hdr = f.readline().decode().strip()
        if hdr not in ("PF", "Pf"):
            raise ValueError(f"{path}: not a PFM file")
        w, h = [int(v) for v in f.readline().decode().split()]
        scale = float(f.readline().decode().strip())
        nc = 3 if hdr == "PF" else 1
        data = np.frombuffer(f.read(4 * w * h * nc), dtype="<f4" if scale < 0 else ">f4")
    img = data.reshape(h, w, nc)[::-1].astype(np.float32) * abs(scale)
    return img


def write_pfm(path, img):
    img = np.asarray(img, dtype=np.float32)
    h, w = img.shape[:2]
    with open(path, "wb") as f:
        f.write(b"PF\n" if img.shape[2] == 3 else b"Pf\n")
        f.write(f"{w} {h}\n-1.0\n".encode())
        f.write(np.ascontiguousarray(img[::-1]).astype("<f4").tobytes())


def rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)))


def channel_rmse(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return [float(np.sqrt(np.mean((a[..., c] - b[..., c]) ** 2))) for c in range(a.shape[-1])]


def read_raw_channels(path):
    """Multi-channel dump written by the oracle build of the reference for "*.exr" outputs (oracle/build_ref.py):
    header "BPTRAW1 w h nc", nc channel names, float32 pixels.  Returns (names, array [h, w, nc])."""
    with open(path, "rb") as f:
        hdr = f.readline().split()
        assert hdr[0] == b"BPTRAW1", hdr
        w, h, nc = int(hdr[1]), int(hdr[2]), int(hdr[3])
        names = [f.readline().decode().strip() for _ in range(nc)]
        data = np.frombuffer(f.read(), np.float32).reshape(h, w, nc)
    return names, data


def write_exr(path, img, names):
    """[h, w, nc] float32 -> OpenEXR through the engine library's host-side writer (csrc/exr_writer.h)."""
    import ctypes as C
    from . import engine
    lib = engine.load_library()
    img = np.ascontiguousarray(img, np.float32)
    h, w, nc = img.shape
    arr = (C.c_char_p * nc)(*[n.encode() for n in names])
    rc = lib.bpt_write_exr_file(str(path).encode(), w, h, nc, arr, img.ctypes.data)
    if rc != 0:
        raise IOError(f"bpt_write_exr_file({path}) failed: {rc}")


def read_exr(path):
    """Independent reader of the subset of OpenEXR the writer emits (single part, scan lines, NO_COMPRESSION, FLOAT / HALF / UINT
    channels): returns (names, array [h, w, nc] float32, attributes).  Written from the published file-format description, not from
    the writer, so that the tests check the file and not a round trip through shared code."""
    import struct
    b = open(path, "rb").read()
    magic, version = struct.unpack_from("<ii", b, 0)
    assert magic == 20000630 and (version & 0xff) == 2 and (version & ~0xff) == 0, (magic, version)
    pos, attrs = 8, {}
    def cstr(p):
        e = b.index(b"\0", p)
        return b[p:e].decode(), e + 1
    while b[pos] != 0:
        name, pos = cstr(pos); typ, pos = cstr(pos)
        size, = struct.unpack_from("<i", b, pos); pos += 4
        attrs[name] = (typ, b[pos:pos + size]); pos += size
    pos += 1
    assert attrs["compression"][1] == b"\0" and attrs["lineOrder"][1] == b"\0"
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    chans, p, raw = [], 0, attrs["channels"][1]
    while raw[p] != 0:
        e = raw.index(b"\0", p); nm = raw[p:e].decode(); p = e + 1
        ptype, = struct.unpack_from("<i", raw, p); p += 16
        chans.append((nm, ptype))
    assert [c[0] for c in chans] == sorted(c[0] for c in chans), "channels must be stored in alphabetical order"
    offs = struct.unpack_from("<%dQ" % h, b, pos)
    out = np.zeros((h, w, len(chans)), np.float32)
    for row in range(h):
        y, nbytes = struct.unpack_from("<ii", b, offs[row]); q = offs[row] + 8
        for k, (nm, ptype) in enumerate(chans):
            dt, sz = {0: ("<u4", 4), 1: ("<f2", 2), 2: ("<f4", 4)}[ptype]
            out[y - y0, :, k] = np.frombuffer(b, dt, w, q).astype(np.float32); q += sz * w
        assert q - offs[row] - 8 == nbytes
    return [c[0] for c in chans], out, attrs
