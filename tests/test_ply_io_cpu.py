"""f4: the explicit-Gaussian and anchor PLY layouts (scene/lod_model.py:374-465,681-832; merge.py:42-53,205-217)."""
import numpy as np
import pytest

from horizongs_b200 import ply_io as P


def test_explicit_attribute_order_matches_reference_lists():
    # merge.py:42-53 for max_sh_degree = 2 (K = 9): 3 DC + 24 rest
    names = P.explicit_attribute_names(9)
    assert names[:8] == ["x", "y", "z", "level", "extra_level", "f_dc_0", "f_dc_1", "f_dc_2"]
    assert names[8] == "f_rest_0" and names[8 + 23] == "f_rest_23" and names[32] == "opacity"
    assert names[33:36] == ["scale_0", "scale_1", "scale_2"] and names[36:] == ["rot_0", "rot_1", "rot_2", "rot_3"]
    a = P.anchor_attribute_names(10, 32)
    assert a[5] == "f_offset_0" and a[34] == "f_offset_29" and a[35] == "f_anchor_feat_0" and a[66] == "f_anchor_feat_31"
    assert a[67:73] == [f"scale_{i}" for i in range(6)] and a[73:] == [f"rot_{i}" for i in range(4)]


def test_explicit_round_trip_and_channel_major_sh(tmp_path):
    g = np.random.default_rng(0)
    N, K = 257, 9
    d = dict(xyz=g.normal(size=(N, 3)), level=g.integers(0, 8, N), extra_level=g.normal(size=N), sh=g.normal(size=(N, K, 3)),
             opacity=g.random(N), scales=g.random((N, 3)), rots=g.normal(size=(N, 4)))
    path = str(tmp_path / "pc" / "point_cloud_explicit.ply")
    P.save_explicit_gaussians(path, d["xyz"], d["level"], d["extra_level"], d["sh"], d["opacity"], d["scales"], d["rots"],
                              26.686, 3, 8)
    names, raw, info = P.read_ply(path)
    assert names == P.explicit_attribute_names(K) and raw.shape == (N, 5 + 3 * K + 1 + 3 + 4)
    assert info == {"standard_dist": pytest.approx(26.686), "aerial_levels": 3.0, "street_levels": 8.0}
    # channel-major storage: f_rest_0..7 are the red coefficients 1..8, f_rest_8.. the green ones (lod_model.py:762)
    assert np.allclose(raw[:, names.index("f_rest_0")], d["sh"][:, 1, 0].astype(np.float32))
    assert np.allclose(raw[:, names.index("f_rest_8")], d["sh"][:, 1, 1].astype(np.float32))
    assert np.allclose(raw[:, names.index("f_dc_2")], d["sh"][:, 0, 2].astype(np.float32))
    out = P.load_explicit_gaussians(path)
    assert out["colors"].shape == (N, K, 3) and np.array_equal(out["colors"], d["sh"].astype(np.float32))
    assert np.array_equal(out["xyz"], d["xyz"].astype(np.float32)) and np.array_equal(out["level"], d["level"].astype(np.int16))
    assert np.array_equal(out["rots"], d["rots"].astype(np.float32)) and out["aerial_levels"] == 3 and out["street_levels"] == 8
    # header exactly as plyfile lays it out
    head = open(path, "rb").read(400).split(b"end_header")[0].decode().splitlines()
    assert head[:5] == ["ply", "format binary_little_endian 1.0", "obj_info standard_dist 26.686000",
                        "obj_info aerial_levels 3.000000", "obj_info street_levels 8.000000"]
    assert head[5] == f"element vertex {N}" and head[6] == "property float x"


def test_anchor_round_trip_offsets_transposed(tmp_path):
    g = np.random.default_rng(1)
    A, k, F = 100, 10, 32
    off = g.normal(size=(A, k, 3))
    path = str(tmp_path / "point_cloud.ply")
    P.save_anchors(path, g.normal(size=(A, 3)), g.integers(0, 5, A), g.normal(size=A), off, g.normal(size=(A, F)),
                   g.normal(size=(A, 6)), np.tile([1.0, 0, 0, 0], (A, 1)), 26.686, 3, 8)
    names, raw, _ = P.read_ply(path)
    # stored [A,3,k]: f_offset_0..9 are the x components of the k offsets (lod_model.py:399)
    assert np.allclose(raw[:, names.index("f_offset_3")], off[:, 3, 0].astype(np.float32))
    assert np.allclose(raw[:, names.index("f_offset_13")], off[:, 3, 1].astype(np.float32))
    out = P.load_anchors(path)
    assert np.array_equal(out["offset"], off.astype(np.float32)) and out["anchor_feat"].shape == (A, F)
    assert out["scaling"].shape == (A, 6) and np.array_equal(out["rotation"][:, 0], np.ones(A, np.float32))


def test_reader_rejects_what_it_does_not_support(tmp_path):
    p = tmp_path / "bad.ply"
    p.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0.0\n")
    with pytest.raises(ValueError):
        P.read_ply(str(p))
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\ncomment made by hand\nelement vertex 2\nproperty float x\n"
                  b"property uchar red\nend_header\n")
    with pytest.raises(ValueError):
        P.read_ply(str(p))
    p.write_bytes(b"ply\nformat binary_little_endian 1.0\ncomment ok\nelement vertex 2\nproperty float32 x\nend_header\n"
                  + np.array([1.5, -2.0], "<f4").tobytes())
    names, data, info = P.read_ply(str(p))
    assert names == ["x"] and data[:, 0].tolist() == [1.5, -2.0] and info == {}


# ---- fixtures in the reference's plyfile layout (tests/golden/make_golden_ply.py, written without ply_io) --------------
def _golden(name):
    import os
    return os.path.join(os.path.dirname(__file__), "golden", name)


def test_reads_anchor_ply_in_the_reference_layout_and_rewrites_it_byte_for_byte(tmp_path):
    """scene/lod_model.py:374-418 save_ply layout: values come back in the model's layout (load_ply, :420-465) and
    save_anchors() reproduces the file exactly"""
    import numpy as np
    from horizongs_b200 import ply_io
    G = np.load(_golden("reference_layout_ply.npz"))
    got = ply_io.load_anchors(_golden("reference_layout_anchor.ply"))
    assert np.array_equal(got["anchor"], G["anchor"]) and np.array_equal(got["offset"], G["offset"])
    assert np.array_equal(got["anchor_feat"], G["feat"]) and np.array_equal(got["scaling"], G["scaling"])
    assert np.array_equal(got["rotation"], G["rot"]) and np.array_equal(got["extra_level"], G["extra"][:, 0])
    assert got["level"].dtype == np.int16 and np.array_equal(got["level"], G["level"][:, 0].astype(np.int16))
    assert got["aerial_levels"] == 3 and got["street_levels"] == 8 and abs(got["standard_dist"] - 26.686) < 1e-6
    out = str(tmp_path / "a.ply")
    ply_io.save_anchors(out, got["anchor"], got["level"], got["extra_level"], got["offset"], got["anchor_feat"],
                        got["scaling"], got["rotation"], got["standard_dist"], got["aerial_levels"], got["street_levels"])
    assert open(out, "rb").read() == open(_golden("reference_layout_anchor.ply"), "rb").read()


def test_reads_explicit_ply_in_the_reference_layout_and_rewrites_it_byte_for_byte(tmp_path):
    """scene/lod_model.py:681-779 save_explicit layout (channel-major SH): colours come back as [N,K,3] (what
    generate_explicit_gaussians feeds the rasterizer, basic_model.py:373-383), and the writer reproduces the file"""
    import numpy as np
    from horizongs_b200 import ply_io
    G = np.load(_golden("reference_layout_ply.npz"))
    got = ply_io.load_explicit_gaussians(_golden("reference_layout_explicit.ply"))
    assert np.array_equal(got["xyz"], G["xyz"]) and np.array_equal(got["colors"], G["color"])
    assert np.array_equal(got["opacity"], G["opacity"][:, 0]) and np.array_equal(got["scales"], G["scale"])
    assert np.array_equal(got["rots"], G["rotation"]) and np.array_equal(got["level"], G["lvl"][:, 0].astype(np.int16))
    out = str(tmp_path / "e.ply")
    ply_io.save_explicit_gaussians(out, got["xyz"], got["level"], got["extra_level"], got["colors"], got["opacity"],
                                   got["scales"], got["rots"], got["standard_dist"], got["aerial_levels"],
                                   got["street_levels"])
    assert open(out, "rb").read() == open(_golden("reference_layout_explicit.ply"), "rb").read()
