"""On-disk formats of the reference's map outputs (host I/O, no arithmetic).

* ``global_map_offline.pcd`` -- ``o3d.io.write_point_cloud(Config.OUTPUT_PCD, final_map)``
  (duc/ICP_LIDAR/slam_offline.py:445-452): PCD v0.7, ``FIELDS x y z``, float32, ``DATA binary``
  (header layout: the bundled global_map_offline.pcd:1-11).
* ``realtime_occupancy_map.png`` -- ``cv2.imwrite(Config.OUTPUT_OCCUPANCY_MAP, occupancy_map)``
  (slam_offline.py:453): 8-bit RGB PNG of the (h, w, 3) BGR picture.

Written with the standard library only (``zlib``, ``struct``), so the files can be produced on a
box without Open3D / OpenCV and read back by them.
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

_PCD_HEADER = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\n"
               "TYPE F F F\nCOUNT 1 1 1\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\n"
               "DATA binary\n")


def write_pcd(path: str, points) -> None:
    """Points (N, 2) or (N, 3) -> binary float32 xyz PCD, byte-compatible with Open3D's writer for
    a cloud that has only points (global_map_offline.pcd)."""
    p = np.asarray(points)
    if p.ndim != 2 or p.shape[1] not in (2, 3):
        raise ValueError("points must be (N, 2) or (N, 3)")
    xyz = np.zeros((len(p), 3), dtype="<f4")
    xyz[:, :p.shape[1]] = p
    with open(path, "wb") as f:
        f.write(_PCD_HEADER.format(n=len(p)).encode("ascii"))
        f.write(xyz.tobytes())


def read_pcd(path: str) -> np.ndarray:
    """Binary / ascii ``x y z`` float32 PCD -> (N, 3) float32."""
    with open(path, "rb") as f:
        blob = f.read()
    fields, n, pos, kind = None, None, 0, None
    while kind is None:
        end = blob.index(b"\n", pos)
        line = blob[pos:end].decode("ascii").strip()
        pos = end + 1
        key, _, rest = line.partition(" ")
        if key == "FIELDS":
            fields = rest.split()
        elif key in ("SIZE", "TYPE", "COUNT"):
            want = {"SIZE": "4", "TYPE": "F", "COUNT": "1"}[key]
            if any(v != want for v in rest.split()):
                raise ValueError(f"unsupported PCD {key}: {rest}")
        elif key == "POINTS":
            n = int(rest)
        elif key == "DATA":
            kind = rest
    if fields is None or n is None or fields[:3] != ["x", "y", "z"]:
        raise ValueError("not an x y z point cloud")
    if kind == "binary":
        a = np.frombuffer(blob, dtype="<f4", count=n * len(fields), offset=pos).reshape(n, len(fields))
    elif kind == "ascii":
        a = np.array(blob[pos:].split(), dtype=np.float32).reshape(n, len(fields))
    else:
        raise ValueError(f"unsupported PCD DATA {kind}")
    return np.ascontiguousarray(a[:, :3], dtype=np.float32)


def _chunk(tag: bytes, data: bytes) -> bytes:
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))


def write_png(path: str, image_bgr, level: int = 6) -> None:
    """(h, w, 3) uint8 BGR picture (OpenCV's channel order) or (h, w) grey -> 8-bit PNG."""
    im = np.asarray(image_bgr)
    if im.dtype != np.uint8 or im.ndim not in (2, 3) or (im.ndim == 3 and im.shape[2] != 3):
        raise ValueError("image must be (h, w, 3) or (h, w) uint8")
    h, w = im.shape[:2]
    rows = im[:, :, ::-1].reshape(h, w * 3) if im.ndim == 3 else im          # BGR -> RGB
    raw = np.zeros((h, rows.shape[1] + 1), dtype=np.uint8)                   # filter byte 0 per row
    raw[:, 1:] = rows
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2 if im.ndim == 3 else 0, 0, 0, 0)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", ihdr) + _chunk(b"IDAT", zlib.compress(raw.tobytes(), level))
                + _chunk(b"IEND", b""))


def read_png(path: str) -> np.ndarray:
    """8-bit grey / RGB non-interlaced PNG -> (h, w) or (h, w, 3) BGR uint8 (as ``cv2.imread``)."""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, hdr = 8, [], None
    while pos < len(blob):
        (ln,), tag = struct.unpack(">I", blob[pos:pos + 4]), blob[pos + 4:pos + 8]
        data = blob[pos + 8:pos + 8 + ln]
        pos += 12 + ln
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", data)
        elif tag == b"IDAT":
            idat.append(data)
        elif tag == b"IEND":
            break
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 8 or ctype not in (0, 2) or interlace:
        raise ValueError("only 8-bit grey / RGB non-interlaced PNG")
    bpp = 3 if ctype == 2 else 1
    raw = np.frombuffer(zlib.decompress(b"".join(idat)), dtype=np.uint8).reshape(h, w * bpp + 1)
    out = np.zeros((h, w * bpp), dtype=np.uint8)
    prev = np.zeros(w * bpp, dtype=np.int32)
    for y in range(h):
        ft, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        if ft == 0:
            cur = line
        elif ft == 2:
            cur = (line + prev) & 255
        else:                                   # Sub / Average / Paeth need the left neighbour
            cur = np.zeros_like(line)
            for i in range(len(line)):
                a = cur[i - bpp] if i >= bpp else 0
                b = prev[i]
                c = prev[i - bpp] if i >= bpp else 0
                if ft == 1:
                    pred = a
                elif ft == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if pa <= pb and pa <= pc else (b if pb <= pc else c)
                cur[i] = (line[i] + pred) & 255
        out[y] = cur
        prev = cur
    return out.reshape(h, w, 3)[:, :, ::-1].copy() if bpp == 3 else out
