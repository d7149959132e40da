"""Golden PLY fixtures in the layout the reference writes with `plyfile`.

`plyfile` is not installed here, so the files are produced by restating, with numpy only, exactly what the reference
does and what plyfile's writer emits for it -- independently of horizongs_b200/ply_io.py (this script does not import
it):
  * the record dtype is `[(attribute, 'f4') for attribute in construct_list_of_attributes()]`, filled with
    `elements[:] = list(map(tuple, attributes))` (scene/lod_model.py:406-411 anchors, :766-770 explicit Gaussians);
  * PlyData([el], obj_info=[...]).write() emits: "ply", "format binary_little_endian 1.0", one "obj_info <text>" line
    per entry, "element vertex <N>", one "property float <name>" per field ('f4' -> "float"), "end_header", then
    the packed little-endian records (lod_model.py:412-418, :771-779).
Attribute order: lod_model.py:375-391 (anchors), merge.py:42-53 + the x,y,z,level,extra_level prefix of
lod_model.py:681-699 (explicit).  Run:  python tests/golden/make_golden_ply.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def plyfile_write(path, names, attributes, obj_info):
    dtype_full = [(n, "f4") for n in names]
    elements = np.empty(attributes.shape[0], dtype=dtype_full)
    elements[:] = list(map(tuple, attributes))
    header = ["ply", "format binary_little_endian 1.0"] + [f"obj_info {t}" for t in obj_info]
    header += [f"element vertex {len(elements)}"] + [f"property float {n}" for n in names] + ["end_header"]
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(elements.astype(elements.dtype.newbyteorder("<")).tobytes())


def main():
    rng = np.random.default_rng(20240607)
    info = dict(standard_dist=26.686, aerial_levels=3, street_levels=8)
    obj_info = ["standard_dist {:.6f}".format(info["standard_dist"]), "aerial_levels {:.6f}".format(info["aerial_levels"]),
                "street_levels {:.6f}".format(info["street_levels"])]
    # ---- anchors (save_ply, lod_model.py:374-418): _offset [A,k,3] stored transposed, _scaling [A,6], _rotation [A,4]
    A, k, F = 7, 10, 32
    anchor = rng.normal(size=(A, 3)).astype(np.float32)
    level = rng.integers(0, 8, size=(A, 1)).astype(np.float32)
    extra = rng.normal(size=(A, 1)).astype(np.float32)
    offset = rng.normal(size=(A, k, 3)).astype(np.float32)
    feat = rng.normal(size=(A, F)).astype(np.float32)
    scaling = rng.normal(size=(A, 6)).astype(np.float32)
    rot = rng.normal(size=(A, 4)).astype(np.float32)
    names = ["x", "y", "z", "level", "extra_level"] + [f"f_offset_{i}" for i in range(k * 3)]
    names += [f"f_anchor_feat_{i}" for i in range(F)] + [f"scale_{i}" for i in range(6)] + [f"rot_{i}" for i in range(4)]
    offsets_t = np.ascontiguousarray(offset.transpose(0, 2, 1)).reshape(A, -1)       # transpose(1, 2).flatten(1)
    attributes = np.concatenate((anchor, level, extra, offsets_t, feat, scaling, rot), axis=1)
    plyfile_write(os.path.join(HERE, "reference_layout_anchor.ply"), names, attributes, obj_info)
    # ---- explicit Gaussians (save_explicit, lod_model.py:681-779): colour [N,K,3] -> f_dc / f_rest channel-major
    N, K = 11, 9
    xyz = rng.normal(size=(N, 3)).astype(np.float32)
    lvl = rng.integers(0, 8, size=(N, 1)).astype(np.float32)
    ext = rng.normal(size=(N, 1)).astype(np.float32)
    color = rng.normal(size=(N, K, 3)).astype(np.float32)
    opacity = rng.uniform(size=(N, 1)).astype(np.float32)
    scale = rng.uniform(size=(N, 3)).astype(np.float32)
    rotation = rng.normal(size=(N, 4)).astype(np.float32)
    f_dc = np.ascontiguousarray(color[:, 0:1, :].transpose(0, 2, 1)).reshape(N, -1)
    f_rest = np.ascontiguousarray(color[:, 1:, :].transpose(0, 2, 1)).reshape(N, -1)
    enames = ["x", "y", "z", "level", "extra_level"] + [f"f_dc_{i}" for i in range(3)]
    enames += [f"f_rest_{i}" for i in range(3 * K - 3)] + ["opacity"] + [f"scale_{i}" for i in range(3)]
    enames += [f"rot_{i}" for i in range(4)]
    eattr = np.concatenate((xyz, lvl, ext, f_dc, f_rest, opacity, scale, rotation), axis=1)
    plyfile_write(os.path.join(HERE, "reference_layout_explicit.ply"), enames, eattr, obj_info)
    np.savez(os.path.join(HERE, "reference_layout_ply.npz"), anchor=anchor, level=level, extra=extra, offset=offset,
             feat=feat, scaling=scaling, rot=rot, xyz=xyz, lvl=lvl, ext=ext, color=color, opacity=opacity, scale=scale,
             rotation=rotation, **{f"info_{k_}": v for k_, v in info.items()})
    print("wrote reference_layout_anchor.ply, reference_layout_explicit.ply, reference_layout_ply.npz")


if __name__ == "__main__":
    main()
