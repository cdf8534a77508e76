"""Wire formats either side of the hot path.

  scene vertices   the reference reads `<scene>.ply` with open3d (fitting_habitat.py:93-96);
                   open3d is absent here, so PLY vertex positions are parsed directly
                   (ascii and binary_little_endian); `.npy` [M,3] is accepted as well.
  body pickles     keys transl, global_orient, betas, body_pose (32-D VPoser latent),
                   left_hand_pose, right_hand_pose, cam_ext, cam_int
                   (test_habitat_s2.py:215-229, cvae.py:305-334).
"""
from __future__ import annotations

import pickle

import numpy as np

_PLY_TYPES = {"char": "i1", "uchar": "u1", "short": "i2", "ushort": "u2", "int": "i4", "uint": "u4",
              "float": "f4", "double": "f8", "int8": "i1", "uint8": "u1", "int16": "i2",
              "uint16": "u2", "int32": "i4", "uint32": "u4", "float32": "f4", "float64": "f8"}


def read_ply_vertices(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, nvert, props, in_vertex = None, 0, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    nvert = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError("list property inside the vertex element")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        names = [p[0] for p in props]
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=nvert, ndmin=2)
            cols = [names.index(c) for c in ("x", "y", "z")]
            return np.ascontiguousarray(rows[:, cols], dtype=np.float32)
        if fmt == "binary_little_endian":
            dt = np.dtype([(n, "<" + t) for n, t in props])
            data = np.frombuffer(f.read(dt.itemsize * nvert), dtype=dt, count=nvert)
            return np.ascontiguousarray(np.stack([data["x"], data["y"], data["z"]], axis=1), dtype=np.float32)
        raise ValueError(f"{path}: unsupported PLY format {fmt}")


def write_ply_vertices(path: str, verts: np.ndarray, binary: bool = True) -> None:
    verts = np.asarray(verts, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(b"ply\n")
        f.write(b"format binary_little_endian 1.0\n" if binary else b"format ascii 1.0\n")
        f.write(f"element vertex {len(verts)}\n".encode())
        f.write(b"property float x\nproperty float y\nproperty float z\nend_header\n")
        if binary:
            f.write(verts.astype("<f4").tobytes())
        else:
            for v in verts:
                f.write(("%.9g %.9g %.9g\n" % tuple(v)).encode())


def read_scene_vertices(path: str) -> np.ndarray:
    if path.endswith(".npy"):
        return np.ascontiguousarray(np.load(path), dtype=np.float32).reshape(-1, 3)
    return read_ply_vertices(path)


BODY_KEYS = ("transl", "global_orient", "betas", "body_pose", "left_hand_pose", "right_hand_pose")


def write_body_pickle(path: str, xh_row: np.ndarray, cam_ext: np.ndarray, cam_int: np.ndarray) -> None:
    """One generated / fitted body in the reference's pickle layout."""
    x = np.asarray(xh_row, dtype=np.float32).reshape(1, 72)
    d = {"transl": x[:, 0:3], "global_orient": x[:, 3:6], "betas": x[:, 6:16], "body_pose": x[:, 16:48],
         "left_hand_pose": x[:, 48:60], "right_hand_pose": x[:, 60:72],
         "cam_ext": np.asarray(cam_ext, dtype=np.float32), "cam_int": np.asarray(cam_int, dtype=np.float32)}
    with open(path, "wb") as f:
        pickle.dump(d, f)


def read_body_pickle(path: str) -> dict:
    with open(path, "rb") as f:
        return pickle.load(f)
