"""On-disk formats either side of the rasterization path: the two PLY layouts
Horizon-GS writes with `plyfile` -- binary little-endian, one float32 `vertex` element, three `obj_info` lines.

  * explicit Gaussians ("point_cloud_explicit.ply"): scene/lod_model.py:681-783 save_explicit, :785-832 load_explicit,
    merge.py:42-53,205-217.  Field order
        x y z level extra_level f_dc_0..2 f_rest_0..(3K-4) opacity scale_0..2 rot_0..3
    with the SH coefficients stored CHANNEL-MAJOR (features.transpose(1, 2).flatten(1), lod_model.py:761-762) and
    transposed back on load (:822-823).  `load_explicit_gaussians` returns them in the rasterizer's [N,K,3] layout
    (what generate_explicit_gaussians concatenates, scene/basic_model.py:373-383).
  * anchors ("point_cloud.ply"): scene/lod_model.py:374-418 save_ply, :420-465 load_ply.  Field order
        x y z level extra_level f_offset_0..(3k-1) f_anchor_feat_* scale_0..5 rot_0..3
    with the offsets stored as [A,3,k] (transpose(1, 2).flatten(1)) and returned as [A,k,3].

Host-side numpy only (no plyfile dependency: it is not installed here); the reader accepts what plyfile writes
(`property float name`, comments, obj_info) and nothing more exotic.  Not a product storage engine: just enough to
feed real Block_A outputs to `rasterization()` instead of synthetic scenes."""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

_FLOAT_TYPES = ("float", "float32")


def write_ply(path: str, names: Sequence[str], data: np.ndarray, obj_info: Dict[str, float]) -> None:
    """one `vertex` element of float32 properties `names` (data [N, len(names)]), binary little-endian, with
    `obj_info <key> <value:.6f>` header lines in the order given (plyfile's layout: format, obj_info, element)."""
    data = np.ascontiguousarray(data, dtype="<f4")
    assert data.ndim == 2 and data.shape[1] == len(names), (data.shape, len(names))
    head = ["ply", "format binary_little_endian 1.0"]
    head += [f"obj_info {k} {float(v):.6f}" for k, v in obj_info.items()]
    head += [f"element vertex {data.shape[0]}"]
    head += [f"property float {n}" for n in names]
    head += ["end_header"]
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(("\n".join(head) + "\n").encode("ascii"))
        f.write(data.tobytes())


def read_ply(path: str) -> Tuple[List[str], np.ndarray, Dict[str, float]]:
    """-> (property names, data [N, F] float32, obj_info) of a float-only binary little-endian vertex element"""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        names: List[str] = []
        info: Dict[str, float] = {}
        n, fmt, in_vertex = None, None, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").strip().split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "obj_info" and len(tok) >= 3:
                info[tok[1]] = float(tok[2])
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
                elif int(tok[2]) != 0:
                    raise ValueError(f"{path}: unsupported non-empty element '{tok[1]}'")
            elif tok[0] == "property" and in_vertex:
                if tok[1] not in _FLOAT_TYPES:
                    raise ValueError(f"{path}: unsupported property type '{' '.join(tok[1:])}' (float32 only)")
                names.append(tok[2])
            elif tok[0] == "end_header":
                break
        if fmt != "binary_little_endian" or n is None:
            raise ValueError(f"{path}: only binary_little_endian PLY with a vertex element is supported")
        raw = f.read(n * len(names) * 4)
        if len(raw) != n * len(names) * 4:
            raise ValueError(f"{path}: truncated PLY body")
    return names, np.frombuffer(raw, dtype="<f4").reshape(n, len(names)).astype(np.float32), info


def _cols(names: List[str], data: np.ndarray, prefix: str) -> np.ndarray:
    """columns whose name starts with prefix, ordered by their trailing integer (the reference's sorted(...) idiom)"""
    sel = sorted((nm for nm in names if nm.startswith(prefix)), key=lambda x: int(x.split("_")[-1]))
    return data[:, [names.index(nm) for nm in sel]]


def _levels(obj_info: Dict[str, float]) -> Dict[str, float]:
    out = dict(obj_info)
    for k in ("aerial_levels", "street_levels"):
        if k in out:
            out[k] = int(round(out[k]))          # lod_model.py:462-463,829-830
    return out


# ---------------------------------------------------------------------------------------- explicit Gaussians
def explicit_attribute_names(n_sh: int) -> List[str]:
    """merge.py:42-53 with the ['x','y','z','level','extra_level'] prefix of merge.py:205; n_sh = (max_sh_degree+1)^2"""
    names = ["x", "y", "z", "level", "extra_level"] + [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(3 * n_sh - 3)]
    return names + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def save_explicit_gaussians(path, xyz, level, extra_level, sh_coeffs, opacity, scales, rots, standard_dist,
                            aerial_levels, street_levels) -> None:
    """sh_coeffs [N,K,3] (rasterizer layout); level / extra_level / opacity [N] or [N,1]"""
    xyz = np.asarray(xyz, np.float32)
    N = xyz.shape[0]
    sh = np.asarray(sh_coeffs, np.float32).reshape(N, -1, 3)
    f_dc = sh[:, 0:1, :].transpose(0, 2, 1).reshape(N, -1)          # channel-major, lod_model.py:761
    f_rest = sh[:, 1:, :].transpose(0, 2, 1).reshape(N, -1)         # :762
    col = lambda a: np.asarray(a, np.float32).reshape(N, -1)         # noqa: E731
    data = np.concatenate([xyz, col(level), col(extra_level), f_dc, f_rest, col(opacity), col(scales), col(rots)], 1)
    write_ply(path, explicit_attribute_names(sh.shape[1]), data,
              {"standard_dist": standard_dist, "aerial_levels": aerial_levels, "street_levels": street_levels})


def load_explicit_gaussians(path) -> Dict[str, object]:
    """-> xyz [N,3], level [N] int16, extra_level [N], colors [N,K,3] (DC first), opacity [N], scales [N,3],
    rots [N,4] (wxyz), plus the obj_info scalars (standard_dist, aerial_levels, street_levels)"""
    names, data, info = read_ply(path)
    N = data.shape[0]
    ix = names.index
    f_dc = np.stack([data[:, ix(f"f_dc_{i}")] for i in range(3)], 1)[:, :, None]        # [N,3,1]
    f_rest = _cols(names, data, "f_rest_").reshape(N, 3, -1)                              # [N,3,K-1], :817
    colors = np.concatenate([f_dc, f_rest], 2).transpose(0, 2, 1).copy()                  # [N,K,3], :822-823
    out = {"xyz": data[:, [ix("x"), ix("y"), ix("z")]].copy(), "level": data[:, ix("level")].astype(np.int16),
           "extra_level": data[:, ix("extra_level")].copy(), "colors": colors, "opacity": data[:, ix("opacity")].copy(),
           "scales": _cols(names, data, "scale_").copy(), "rots": _cols(names, data, "rot").copy()}
    out.update(_levels(info))
    return out


# ---------------------------------------------------------------------------------------- anchors
def anchor_attribute_names(n_offsets: int, feat_dim: int, n_scale: int = 6, n_rot: int = 4) -> List[str]:
    """scene/lod_model.py:375-391 construct_list_of_attributes"""
    names = ["x", "y", "z", "level", "extra_level"] + [f"f_offset_{i}" for i in range(3 * n_offsets)]
    names += [f"f_anchor_feat_{i}" for i in range(feat_dim)]
    return names + [f"scale_{i}" for i in range(n_scale)] + [f"rot_{i}" for i in range(n_rot)]


def save_anchors(path, anchor, level, extra_level, offset, anchor_feat, scaling, rotation, standard_dist,
                 aerial_levels, street_levels) -> None:
    """offset [A,k,3] (model layout), stored as [A,3,k] flattened (lod_model.py:399)"""
    anchor = np.asarray(anchor, np.float32)
    A = anchor.shape[0]
    off = np.asarray(offset, np.float32).reshape(A, -1, 3)
    col = lambda a: np.asarray(a, np.float32).reshape(A, -1)         # noqa: E731
    data = np.concatenate([anchor, col(level), col(extra_level), off.transpose(0, 2, 1).reshape(A, -1), col(anchor_feat),
                           col(scaling), col(rotation)], 1)
    names = anchor_attribute_names(off.shape[1], col(anchor_feat).shape[1], col(scaling).shape[1], col(rotation).shape[1])
    write_ply(path, names, data,
              {"standard_dist": standard_dist, "aerial_levels": aerial_levels, "street_levels": street_levels})


def load_anchors(path) -> Dict[str, object]:
    """-> anchor [A,3], level [A] int16, extra_level [A], offset [A,k,3], anchor_feat [A,F], scaling [A,6],
    rotation [A,4], plus the obj_info scalars"""
    names, data, info = read_ply(path)
    A = data.shape[0]
    ix = names.index
    off = _cols(names, data, "f_offset").reshape(A, 3, -1).transpose(0, 2, 1).copy()     # :447,452
    out = {"anchor": data[:, [ix("x"), ix("y"), ix("z")]].copy(), "level": data[:, ix("level")].astype(np.int16),
           "extra_level": data[:, ix("extra_level")].copy(), "offset": off,
           "anchor_feat": _cols(names, data, "f_anchor_feat").copy(), "scaling": _cols(names, data, "scale_").copy(),
           "rotation": _cols(names, data, "rot").copy()}
    out.update(_levels(info))
    return out
