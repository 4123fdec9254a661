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
