"""On-disk formats on either side of the PointFlow path.

Same names, argument meaning and FILE BYTES as the reference's `pointmvsnet/utils/io.py`
(camera text files `io.py:15-75`, PFM `io.py:78-145`) and the Gipuma `.dmb` container of
`tools/depthfusion.py:27-61`; pinned byte-for-byte against files written by the reference
(`tests/golden/io_golden.npz`, `tests/test_host.py`).  Host-side code: numpy only.
"""
import os
import re
import struct
import sys

import numpy as np

__all__ = ["mkdir", "load_cam_dtu", "write_cam_dtu", "load_pfm", "write_pfm", "read_gipuma_dmb",
           "write_gipuma_dmb"]


def mkdir(path):
    os.makedirs(path, exist_ok=True)


# ----------------------------------------------------------------------------------------
# camera text files: "extrinsic" 4x4, "intrinsic" 3x3, then depth_min interval [num_depth [depth_max]]
# ----------------------------------------------------------------------------------------
def load_cam_dtu(file, num_depth=0, interval_scale=1.0):
    """Parse an MVSNet camera file object into a [2,4,4] float64 array (io.py:15-52).

    cam[0] = extrinsic; cam[1][:3,:3] = intrinsic; cam[1][3] = (depth_min, interval*scale,
    num_depth, depth_max).  The token count decides which of the trailing fields the file holds
    (29: min, interval; 30: + num_depth; 31: + depth_max); any other count leaves the row zero.
    """
    tok = file.read().split()
    cam = np.zeros((2, 4, 4))
    cam[0] = np.array([float(t) for t in tok[1:17]]).reshape(4, 4)       # tok[0] == "extrinsic"
    cam[1, :3, :3] = np.array([float(t) for t in tok[18:27]]).reshape(3, 3)  # tok[17] == "intrinsic"
    n = len(tok)
    if n in (29, 30, 31):
        dmin = float(tok[27])
        interval = float(tok[28]) * interval_scale
        # with 29 tokens the caller's num_depth is used; with 30 the file's count is stored but the
        # far plane is still derived from the caller's num_depth (io.py:37-41)
        stored = float(num_depth) if n == 29 else float(tok[29])
        dmax = float(tok[30]) if n == 31 else dmin + interval * (num_depth - 1)
        cam[1, 3] = (dmin, interval, stored, dmax)
    return cam


def write_cam_dtu(file, cam):
    """Write a [2,4,4] camera in the layout `load_cam_dtu` reads (io.py:55-75); numbers are printed
    with `str()` of the array's scalar type, one trailing blank per number."""
    rows = ["extrinsic"]
    for i in range(4):
        rows.append("".join(str(cam[0][i][j]) + " " for j in range(4)))
    rows.append("")
    rows.append("intrinsic")
    for i in range(3):
        rows.append("".join(str(cam[1][i][j]) + " " for j in range(3)))
    rows.append("")
    rows.append(" ".join(str(cam[1][3][j]) for j in range(4)))
    with open(file, "w") as f:
        f.write("\n".join(rows) + "\n")


# ----------------------------------------------------------------------------------------
# PFM: "Pf" (grey) / "PF" (colour), "<w> <h>", "<scale>" (negative = little endian), rows bottom-up
# ----------------------------------------------------------------------------------------
def load_pfm(file):
    """Read a PFM file -> (float32 array [H,W] or [H,W,3], scale)  (io.py:78-113)."""
    with open(file, "rb") as f:
        magic = f.readline().rstrip().decode("ascii")
        if magic not in ("PF", "Pf"):
            raise Exception("Not a PFM file.")
        dims = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("ascii"))
        if not dims:
            raise Exception("Malformed PFM header.")
        width, height = int(dims.group(1)), int(dims.group(2))
        scale = float(f.readline().decode("ascii").rstrip())
        order = "<" if scale < 0 else ">"
        data = np.fromfile(f, order + "f")
    shape = (height, width, 3) if magic == "PF" else (height, width)
    return np.flipud(np.reshape(data, shape)), abs(scale)


def write_pfm(file, image, scale=1):
    """Write a float32 image as PFM (io.py:116-145): bottom row first, native byte order recorded in
    the sign of the scale line ("%f")."""
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        magic = b"PF\n"
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        magic = b"Pf\n"
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    byteorder = image.dtype.byteorder
    little = byteorder == "<" or (byteorder == "=" and sys.byteorder == "little")
    with open(file, "wb") as f:
        f.write(magic)
        f.write(b"%d %d\n" % (image.shape[1], image.shape[0]))
        f.write(b"%f\n" % (-scale if little else scale))
        f.write(np.ascontiguousarray(np.flipud(image)).tobytes())


# ----------------------------------------------------------------------------------------
# Gipuma .dmb (tools/depthfusion.py:27-61): int32 type=1, h, w, channels, then float32 data
# ----------------------------------------------------------------------------------------
def read_gipuma_dmb(path):
    with open(path, "rb") as f:
        _type, height, width, channels = struct.unpack("<4i", f.read(16))
        data = np.fromfile(f, np.float32)
    # stored with x fastest inside a row, rows inside a channel (Fortran order of [w,h,c])
    return np.transpose(data.reshape((width, height, channels), order="F"), (1, 0, 2)).squeeze()


def write_gipuma_dmb(path, image):
    image = np.asarray(image)
    height, width = image.shape[0], image.shape[1]
    channels = image.shape[2] if image.ndim == 3 else 1
    if image.ndim == 3:
        image = np.transpose(image, (2, 0, 1)).squeeze()  # channel planes, as the reference stores them
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", 1, height, width, channels))
        image.tofile(f)
