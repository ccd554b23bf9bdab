"""Reader/writer for the tagged-array container used by the oracle tools (oracle/recfile.h)."""
from __future__ import annotations

import struct

import numpy as np

_DTYPES = [np.uint8, np.int16, np.int32, np.float32, np.float64, np.uint64]
_MAGIC = b"SDRR0001"


def write(path: str, arrays: dict) -> None:
    with open(path, "wb") as f:
        f.write(_MAGIC)
        for name, a in arrays.items():
            if isinstance(a, (bytes, str)):
                a = np.frombuffer(a.encode() if isinstance(a, str) else a, dtype=np.uint8)
            a = np.ascontiguousarray(a)
            code = next(i for i, d in enumerate(_DTYPES) if a.dtype == d)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<IQ", code, a.size))
            f.write(a.tobytes())


def read(path: str) -> dict:
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != _MAGIC:
            raise ValueError("bad magic in " + path)
        while True:
            hdr = f.read(4)
            if len(hdr) < 4:
                break
            (nl,) = struct.unpack("<I", hdr)
            name = f.read(nl).decode()
            code, count = struct.unpack("<IQ", f.read(12))
            dt = np.dtype(_DTYPES[code])
            out[name] = np.frombuffer(f.read(count * dt.itemsize), dtype=dt).copy()
    return out
