"""Portable float map (.pfm) images as the reference's demo reads and writes them (demo/image.h:196-240 read, :268-295 write):
'PF' = three channels, 'Pf' = one; a negative scale = little-endian floats. read_pfm returns the rows in FILE order, which
is how Image<C>::read fills its buffer (Image::get(uv) then flips the row index, image.h:103-110); write_pfm takes an image
in that buffer order and writes it bottom row first, like Image<C>::write."""
import numpy as np


def read_pfm(path):
    with open(path, "rb") as f:
        kind = f.readline().strip()
        if kind not in (b"PF", b"Pf"):
            raise ValueError("%s: not a PFM file" % path)
        dims = f.readline().split()
        while len(dims) < 2:
            dims += f.readline().split()
        w, h = int(dims[0]), int(dims[1])
        scale = float(f.readline().strip())
        c = 3 if kind == b"PF" else 1
        data = np.frombuffer(f.read(4*w*h*c), dtype="<f4" if scale < 0 else ">f4")
    if data.size != w*h*c:
        raise ValueError("%s: truncated PFM file" % path)
    return data.astype(np.float32).reshape(h, w, c)


def write_pfm(path, image):
    """image: [h, w] or [h, w, 1 or 3] in the buffer order of the reference's Image (row 0 = top)"""
    a = np.asarray(image, np.float32)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    if c not in (1, 3):
        raise ValueError("write_pfm: 1 or 3 channels")
    with open(path, "wb") as f:
        f.write(b"PF\n" if c == 3 else b"Pf\n")
        f.write(("%d %d\n" % (w, h)).encode())
        f.write(b"-1\n")
        f.write(np.ascontiguousarray(a[::-1]).astype("<f4").tobytes())


def luminance(rgb):
    """Image<1>::setFromRGB (image.h:150-154)"""
    rgb = np.asarray(rgb, np.float32)
    return (np.float32(0.299)*rgb[..., 0] + np.float32(0.587)*rgb[..., 1] + np.float32(0.114)*rgb[..., 2]).astype(np.float32)
