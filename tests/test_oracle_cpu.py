"""CPU tests: the oracle against (a) golden vectors generated from the reference's own helpers,
(b) closed-form known answers, (c) its own float64 autograd; plus host-side checks (ABI, constants)."""
import math
import os
import re

import numpy as np
import pytest
import torch

import oracle
from oracle import constants as K
from oracle import gsplat_oracle as O
from tests.helpers import small_scene

GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_conventions.npz")


# ---------------------------------------------------------------- golden vectors (reference-pinned)
def test_sh_matches_reference_eval_sh():
    """utils/sh_utils.py:57-112 eval_sh, all degrees 0..4, float64."""
    g = np.load(GOLD)
    dirs = torch.from_numpy(g["dirs"])
    coeffs = torch.from_numpy(g["coeffs"])
    for deg in range(5):
        ours = O.spherical_harmonics(deg, dirs * 3.7, coeffs)   # un-normalised input: normalised inside
        ref = torch.from_numpy(g[f"sh_deg{deg}"])
        assert torch.allclose(ours, ref, atol=1e-12, rtol=1e-10), deg


def test_quat_rotation_matches_reference_build_rotation():
    """utils/general_utils.py:113-145: wxyz -> R, and L = R S so that Sigma = L L^T."""
    g = np.load(GOLD)
    q = torch.from_numpy(g["quats"])
    s = torch.from_numpy(g["scales"])
    R = torch.stack([torch.stack(r, -1) for r in O._quat_to_rot(q)], -2)
    assert torch.allclose(R, torch.from_numpy(g["rot"]), atol=2e-6)
    L = R * s[:, None, :]
    assert torch.allclose(L, torch.from_numpy(g["scaling_rot"]), atol=2e-6)


def test_world2view_convention_matches_reference():
    """utils/graphics_utils.py:38-49 + scene/cameras.py:91: viewmat = [R^T | T], OpenCV axes; a point
    projects where the reference's matrix says it does."""
    g = np.load(GOLD)
    V = torch.from_numpy(g["world2view"])
    Rc2w, T = g["cam_R"], g["cam_T"]
    assert np.allclose(V[:3, :3].numpy(), Rc2w.T, atol=1e-6) and np.allclose(V[:3, 3].numpy(), T, atol=1e-6)
    p = torch.tensor([[0.2, -0.1, 1.5]])
    Km = torch.tensor([[100.0, 0, 64], [0, 100.0, 48], [0, 0, 1]])
    radii, m2, depth, conics, _ = O.fully_fused_projection(
        p, None, torch.tensor([[1.0, 0, 0, 0]]), torch.full((1, 3), 0.05), V[None], Km[None], 128, 96)
    pc = V[:3, :3] @ p[0] + V[:3, 3]
    assert radii[0, 0] > 0
    assert torch.allclose(depth[0, 0], pc[2])
    assert torch.allclose(m2[0, 0], torch.stack([100 * pc[0] / pc[2] + 64, 100 * pc[1] / pc[2] + 48]), atol=1e-4)


# ---------------------------------------------------------------- known answers
def _one_gaussian(opacity=0.8, s=0.2, z=5.0, f=50.0, W=64, H=64):
    means = torch.tensor([[0.0, 0.0, z]])
    quats = torch.tensor([[1.0, 0, 0, 0]])
    scales = torch.full((1, 3), s)
    Km = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]])[None]
    return means, quats, scales, torch.tensor([opacity]), torch.tensor([[1.0, 0.5, 0.25]]), torch.eye(4)[None], Km, W, H


def test_single_isotropic_gaussian_closed_form():
    means, quats, scales, op, col, V, Km, W, H = _one_gaussian()
    rc, ra, meta = O.rasterization(means, quats, scales, op, col, V, Km, W, H, render_mode="RGB+ED")
    var = (0.2 * 50 / 5) ** 2 + K.EPS2D_DEFAULT
    assert int(meta["radii"][0, 0]) == math.ceil(3 * math.sqrt(var))
    assert torch.allclose(meta["conics"][0, 0], torch.tensor([1 / var, 0.0, 1 / var]), atol=1e-6)
    for (y, x) in [(31, 31), (30, 33), (28, 35), (32, 26)]:
        r2 = (x + 0.5 - 32) ** 2 + (y + 0.5 - 32) ** 2
        a = 0.8 * math.exp(-r2 / (2 * var))
        a = a if a >= K.ALPHA_MIN else 0.0
        assert abs(float(ra[0, y, x, 0]) - a) < 1e-6
        assert torch.allclose(rc[0, y, x, :3], a * col[0], atol=1e-6)
        if a > 0:
            assert abs(float(rc[0, y, x, 3]) - 5.0) < 1e-5          # expected depth = z
    assert float(ra[0, 0, 0, 0]) == 0.0 and float(rc[0, 0, 0, 3]) == 0.0   # empty pixel: depth 0 (train.py:196)


def test_two_gaussians_order_and_termination():
    """front Gaussian occludes; a nearly opaque stack stops before T <= 1e-4."""
    means = torch.tensor([[0.0, 0.0, 4.0], [0.0, 0.0, 2.0]])
    quats = torch.tensor([[1.0, 0, 0, 0]] * 2)
    scales = torch.full((2, 3), 0.3)
    op = torch.tensor([0.9, 0.5])
    col = torch.tensor([[1.0, 0, 0], [0, 1.0, 0]])
    V, Km = torch.eye(4)[None], torch.tensor([[40.0, 0, 16], [0, 40.0, 16], [0, 0, 1]])[None]
    rc, ra, meta = O.rasterization(means, quats, scales, op, col, V, Km, 32, 32)
    assert meta["flatten_ids"][:1].tolist() == [1]               # nearer Gaussian (index 1) sorts first in its tile
    y = x = 16
    def alpha(o, z):
        var = (0.3 * 40 / z) ** 2 + 0.3
        return min(0.999, o * math.exp(-(0.5 ** 2 + 0.5 ** 2) / (2 * var)))
    a1, a0 = alpha(0.5, 2.0), alpha(0.9, 4.0)
    assert torch.allclose(rc[0, y, x], torch.tensor([(1 - a1) * a0, a1, 0.0]), atol=1e-6)
    # termination: 12 near-opaque layers; the second layer would bring T to ~6e-6 <= 1e-4: stop after the first
    n = 12
    means = torch.tensor([[0.0, 0.0, 2.0 + 0.1 * i] for i in range(n)])
    rc, ra, _ = O.rasterization(means, torch.tensor([[1.0, 0, 0, 0]] * n), torch.full((n, 3), 0.5),
                                torch.full((n,), 1.0), torch.rand(n, 3), V, Km, 32, 32)
    a_first = math.exp(-0.5 / (2 * ((0.5 * 40 / 2.0) ** 2 + 0.3)))
    assert a_first < 0.999 and (1 - a_first) ** 2 < 1e-4
    assert abs(float(ra[0, 16, 16, 0]) - a_first) < 1e-6


def test_tile_corner_and_stable_ties():
    """a Gaussian centred on a tile corner touches exactly the 4 surrounding tiles, emitted row-major;
    equal depths keep Gaussian-index order (stable sort)."""
    means2d = torch.tensor([[[32.0, 32.0], [32.0, 32.0], [8.0, 8.0]]])
    radii = torch.tensor([[5, 5, 3]], dtype=torch.int32)
    depths = torch.tensor([[2.0, 2.0, 1.0]])
    tiles, ids, flat = O.isect_tiles(means2d, radii, depths, 16, 4, 4)
    assert tiles.tolist() == [[4, 4, 1]]
    tile_of = ((ids >> 32) & 0xFFFFFFFF).tolist()
    assert tile_of == [0, 5, 5, 6, 6, 9, 9, 10, 10]
    assert flat.tolist() == [2, 0, 1, 0, 1, 0, 1, 0, 1]
    assert (ids & 0xFFFFFFFF).tolist()[1] == int(np.float32(2.0).view(np.int32))
    off = O.isect_offset_encode(ids, 1, 4, 4).flatten().tolist()
    assert off == [0, 1, 1, 1, 1, 1, 3, 5, 5, 5, 7, 9, 9, 9, 9, 9]


def test_culling_cases():
    V, Km = torch.eye(4)[None], torch.tensor([[50.0, 0, 32], [0, 50.0, 32], [0, 0, 1]])[None]
    means = torch.tensor([[0.0, 0, -1.0], [0.0, 0, 0.005], [100.0, 0, 5.0], [0.0, 0, 5.0], [0.0, 0.0, 2e10]])
    q = torch.tensor([[1.0, 0, 0, 0]] * 5)
    radii, m2, d, con, _ = O.fully_fused_projection(means, None, q, torch.full((5, 3), 0.1), V, Km, 64, 64)
    assert (radii[0] > 0).tolist() == [False, False, False, True, False]
    assert float(m2[0, 0].abs().sum() + con[0, 2].abs().sum()) == 0.0
    # radius_clip
    radii2, *_ = O.fully_fused_projection(means, None, q, torch.full((5, 3), 0.1), V, Km, 64, 64, radius_clip=100.0)
    assert int(radii2.sum()) == 0


def test_empty_inputs():
    V, Km = torch.eye(4)[None], torch.tensor([[50.0, 0, 32], [0, 50.0, 32], [0, 0, 1]])[None]
    z = torch.zeros
    rc, ra, meta = O.rasterization(z(0, 3), z(0, 4), z(0, 3), z(0), z(0, 3), V, Km, 40, 24,
                                   backgrounds=torch.tensor([[0.1, 0.2, 0.3]]))
    assert rc.shape == (1, 24, 40, 3) and float(ra.abs().sum()) == 0
    assert torch.allclose(rc[0, 3, 5], torch.tensor([0.1, 0.2, 0.3]))
    assert meta["isect_offsets"].shape == (1, 2, 3) and int(meta["isect_offsets"].abs().sum()) == 0


# ---------------------------------------------------------------- internal consistency
def test_float64_autograd_gradcheck_small():
    sc, V, Ks, W, H = small_scene(n=12, width=32, height=32, scale=0.25, extent=1.0)
    inp = [t.double().requires_grad_() for t in (sc.means, sc.quats, sc.scales, sc.opacities, sc.colors)]

    def f(m, q, s, o, c):
        rc, ra, _ = O.rasterization(m, q, s, o, c, V.double(), Ks.double(), W, H, render_mode="RGB+ED")
        return (rc * torch.linspace(0.5, 1.5, rc.numel(), dtype=torch.float64).reshape(rc.shape)).sum() + ra.sum()

    assert torch.autograd.gradcheck(f, inp, eps=1e-6, atol=1e-5, rtol=1e-3, nondet_tol=0.0)


def test_2dgs_oracle_smoke_and_normals():
    sc, V, Ks, W, H = small_scene(n=300, width=64, height=48, scale=0.2)
    (rc, ra, rn, rnd, rd, rm), meta = O.rasterization_2dgs(
        sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, V, Ks, W, H, render_mode="RGB+ED", distloss=True)
    assert rc.shape == (1, H, W, 4) and rn.shape == (1, H, W, 3) and rnd.shape == (H, W, 3)
    assert rd.shape == (1, H, W, 1) and rm.shape == (1, H, W, 1)
    n = meta["normals"][0][meta["radii"][0] > 0]
    assert torch.allclose(n.norm(dim=-1), torch.ones(n.shape[0]), atol=1e-5)
    # normals face the camera: n . (centre in camera frame) < 0
    pc = (sc.means @ V[0, :3, :3].T + V[0, :3, 3])[meta["radii"][0] > 0]
    assert bool(((n * pc).sum(-1) <= 0).all())
    assert float(ra.max()) <= 1.0 and float(rd.abs().max()) > 0


def test_sh_bases_orthonormal():
    """Monte-Carlo orthonormality of the 25 basis functions (independent of any reference code)."""
    g = torch.Generator().manual_seed(1)
    d = torch.nn.functional.normalize(torch.randn(400_000, 3, generator=g, dtype=torch.float64), dim=-1)
    B = torch.stack(O._sh_bases(4, d), -1)
    G = 4 * math.pi * (B.T @ B) / d.shape[0]
    assert torch.allclose(G, torch.eye(25, dtype=torch.float64), atol=0.02)


# ---------------------------------------------------------------- host side
def test_constants_agree_between_oracle_and_cuda_header():
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "horizongs_b200", "csrc", "hgs_constants.cuh")
    text = open(path).read()
    def val(name):
        m = re.search(rf"#define\s+{name}\s+(.+)", text)
        expr = m.group(1).strip().replace("f", "").replace("(", "").replace(")", "")
        return eval(expr)
    pairs = {"HGS_TILE_SIZE": K.TILE_SIZE, "HGS_ALPHA_MAX": K.ALPHA_MAX, "HGS_ALPHA_MIN": K.ALPHA_MIN,
             "HGS_T_EPS": K.T_EPS, "HGS_RADIUS_SIGMA": K.RADIUS_SIGMA, "HGS_EIG_FLOOR": K.EIG_FLOOR,
             "HGS_FOV_MARGIN": K.FOV_MARGIN, "HGS_ED_ALPHA_FLOOR": K.ED_ALPHA_FLOOR, "HGS_SH_OFFSET": K.SH_OFFSET,
             "HGS_FILTER_INV_SQUARE_2DGS": K.FILTER_INV_SQUARE_2DGS, "HGS_RADIUS_FLOOR_2DGS": K.RADIUS_FLOOR_2DGS,
             "HGS_MEDIAN_T_2DGS": K.MEDIAN_T_2DGS}
    for name, want in pairs.items():
        assert abs(val(name) - want) <= 1e-12 * max(1.0, abs(want)), name


def test_c_abi_library_loads_and_exports_every_declared_symbol(built_lib):
    from horizongs_b200 import _lib
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/hgs_raster.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert built_lib.hgs_abi_version() == 2
    assert built_lib.hgs_status_string(-1).decode().startswith("hgs:")


def test_product_path_has_no_cpu_fallback_and_no_oracle_import():
    import horizongs_b200
    sc, V, Ks, W, H = small_scene(n=10, width=32, height=32)
    with pytest.raises(ValueError):
        horizongs_b200.rasterization(sc.means, sc.quats, sc.scales, sc.opacities, sc.colors, V, Ks, W, H)
    pkg = os.path.dirname(horizongs_b200.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_gsplat_shim_import_surface():
    """render.py:13-14: `import gsplat` and `from gsplat.cuda._wrapper import ...`"""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(__file__))
    code = ("import sys; sys.path[:0]=[%r, %r]; import gsplat; "
            "from gsplat.cuda._wrapper import fully_fused_projection, fully_fused_projection_2dgs; "
            "assert callable(gsplat.rasterization) and callable(gsplat.rasterization_2dgs)") % (
        root, os.path.join(root, "shim"))
    subprocess.run([sys.executable, "-c", code], check=True)


def test_reference_loss_fixtures_are_consistent():
    """tests/golden/reference_losses.npz (generated from the reference's utils/loss_utils.py by
    tests/golden/make_golden_losses.py): loss = (1 - lambda) * l1 + lambda * (1 - ssim), train.py:158-160; the GPU
    test test_fused_l1_ssim_loss_matches_reference_golden checks csrc/loss.cu against these values."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_losses.npz"))
    for case in ("a", "b", "c"):
        lam = float(G[f"{case}_lambda"])
        assert abs(float(G[f"{case}_loss"]) - ((1 - lam) * float(G[f"{case}_l1"]) + lam * (1 - float(G[f"{case}_ssim"])))) < 1e-12
        assert G[f"{case}_grad"].shape == G[f"{case}_img"].shape
        l1 = np.abs(G[f"{case}_img"].astype(np.float64) - G[f"{case}_gt"].astype(np.float64)).mean()
        assert abs(l1 - float(G[f"{case}_l1"])) < 1e-12


def test_ctypes_signatures_match_the_header_prototypes():
    """every function declared in include/hgs_raster.h is bound in horizongs_b200/_lib.py with the same number of
    parameters and compatible kinds (pointer / integer / float), so the ctypes layer cannot drift from the C ABI"""
    import ctypes as C
    from horizongs_b200 import _lib
    text = open(_lib.HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|long long|unsigned long long|const char\*)\s+(hgs_[a-z0-9_]+)\s*\((.*?)\)\s*;", text,
                        flags=re.S)
    assert len(protos) >= 50
    seen = set()
    for name, params in protos:
        seen.add(name)
        assert name in _lib.SIGNATURES, f"{name} is declared in the header but not bound in _lib.SIGNATURES"
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        argtypes = _lib.SIGNATURES[name][1]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
        for p, t in zip(plist, argtypes):
            is_ptr = "*" in p
            if is_ptr:
                assert t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer), (name, p, t)
            elif re.match(r"(const )?float\b", p):
                assert t is C.c_float, (name, p, t)
            else:
                assert t in (C.c_int, C.c_longlong, C.c_size_t, C.c_ulonglong), (name, p, t)
    assert seen == set(_lib.SIGNATURES), sorted(set(_lib.SIGNATURES) ^ seen)


def test_committed_bench_line_carries_the_contract_keys():
    """the final one-GPU bench line committed under profiles/ has every key of the bench contract, its roofline and
    end-to-end blocks are self-consistent, and its parity block is green (guards bench.py's JSON line against silently
    losing a key)"""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_bench_n1_final.json")
    line = json.loads([x for x in open(path) if x.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    assert abs(line["value"] - 1000.0 / line["ms_per_step"]) < 1e-6 * line["value"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and e2e["value"] < line["value"]
    rf = line["roofline"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0.0 < rf["frac"] < 1.0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    assert line["gpu_launches"] > 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    pc = line["parity_check"]
    assert pc["ok"] is True
    for v in pc["per_view"].values():
        assert all(v["integer_stages_bit_exact"].values())
        assert v["image_max_err"] <= pc["tol"]["image_abs"]
        assert max(v["grad_max_rel"].values()) <= pc["tol"]["grad_rel"]
