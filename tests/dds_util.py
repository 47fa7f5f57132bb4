"""Writes small 3-D scalar DDS files for the ingest tests (DDS_HEADER + optional DDS_HEADER_DXT10)."""
import struct

import numpy as np

DXGI = {"r32f": 41, "r16f": 54, "r16un": 56, "r8un": 61}


def write_dds(path, vol, kind="r32f", dx10=True, mips=1):
    """vol: [z][y][x] array in [0, 1] (float); kind: r32f | r16f | r16un | r8un."""
    d, h, w = vol.shape
    raw = {"r32f": lambda v: v.astype("<f4"), "r16f": lambda v: v.astype("<f2"),
           "r16un": lambda v: np.floor(v * 65535 + 0.5).astype("<u2"), "r8un": lambda v: np.floor(v * 255 + 0.5).astype("u1")}[kind](vol)
    bpt = raw.dtype.itemsize
    flags = 0x1 | 0x2 | 0x4 | 0x1000 | 0x800000 | 0x8          # CAPS | HEIGHT | WIDTH | PIXELFORMAT | DEPTH | PITCH
    if dx10:
        pf = struct.pack("<II4sIIIII", 32, 0x4, b"DX10", 0, 0, 0, 0, 0)
    elif kind in ("r32f", "r16f"):
        pf = struct.pack("<IIIIIIII", 32, 0x4, 114 if kind == "r32f" else 111, 0, 0, 0, 0, 0)
    else:
        pf = struct.pack("<IIIIIIII", 32, 0x20000, 0, 8 * bpt, (1 << (8 * bpt)) - 1, 0, 0, 0)   # DDPF_LUMINANCE
    hdr = struct.pack("<4sIIIIIII", b"DDS ", 124, flags, h, w, w * bpt, d, mips) + b"\0" * 44 + pf + \
        struct.pack("<IIIII", 0x1000 | 0x8, 0x200000, 0, 0, 0)                                   # caps: TEXTURE | COMPLEX, caps2: VOLUME
    assert len(hdr) == 128
    with open(path, "wb") as f:
        f.write(hdr)
        if dx10:
            f.write(struct.pack("<IIIII", DXGI[kind], 4, 0, 1, 0))                                # TEXTURE3D
        f.write(raw.tobytes())
        if mips > 1:
            f.write(b"\0" * (raw.nbytes // 8))
    return raw
