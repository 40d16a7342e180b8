"""Reader/writer for the "PGSGDARR" v1 named-array container (see odgi_b200/host/pgsgd_arrays.hpp).

Used for flattened graphs, layouts and golden vectors.  Files ending in ``.gz`` are gzip-compressed
transparently (the committed fixtures under tests/golden/ are).
"""
from __future__ import annotations

import gzip
import io
import struct
from typing import Dict

import numpy as np

_DTYPES = {0: np.uint8, 1: np.uint32, 2: np.uint64, 3: np.float32, 4: np.float64, 5: np.int64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


def read_arrays(path: str) -> Dict[str, np.ndarray]:
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"PGSGDARR":
        raise ValueError(f"{path}: not a PGSGDARR file")
    version, n = struct.unpack_from("<II", buf, 8)
    if version != 1:
        raise ValueError(f"{path}: unsupported PGSGDARR version {version}")
    off = 16
    out: Dict[str, np.ndarray] = {}
    for _ in range(n):
        (nl,) = struct.unpack_from("<I", buf, off)
        off += 4
        name = buf[off:off + nl].decode()
        off += nl
        dt, count = struct.unpack_from("<IQ", buf, off)
        off += 12
        dtype = np.dtype(_DTYPES[dt])
        nbytes = count * dtype.itemsize
        out[name] = np.frombuffer(buf, dtype=dtype, count=count, offset=off).copy()
        off += nbytes
        off += (-off) % 8
    return out


def write_arrays(path: str, arrays: Dict[str, np.ndarray]) -> None:
    bio = io.BytesIO()
    bio.write(b"PGSGDARR")
    bio.write(struct.pack("<II", 1, len(arrays)))
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        if a.dtype not in _CODES:
            raise TypeError(f"{name}: dtype {a.dtype} not supported")
        nb = name.encode()
        bio.write(struct.pack("<I", len(nb)))
        bio.write(nb)
        bio.write(struct.pack("<IQ", _CODES[a.dtype], a.size))
        bio.write(a.tobytes())
        bio.write(b"\0" * ((-bio.tell()) % 8))
    data = bio.getvalue()
    if str(path).endswith(".gz"):
        with gzip.GzipFile(path, "wb", mtime=0) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)
