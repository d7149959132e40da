"""the reference's own adapter -- gaussian_renderer/render.py:16-197, UNMODIFIED -- is
executed against the drop-in.  On the CPU (build container, /root/reference present) the four gsplat names resolve to
the oracle and every key of render()'s return contract (render.py:96-116) is checked; on the GPU box the same file
(installed to baseline/_ref by __graft_entry__.build()) runs against ``shim/gsplat`` (libhgs_raster.so) and is
compared, key by key and gradient by gradient, with its own run on the oracle."""
import copy

import pytest
import torch

from horizongs_b200 import scenes
from tests import lod_harness as LH
from tests import reference_adapter as RA
from tests.helpers import rel_err, small_scene

needs_reference = pytest.mark.skipif(RA.reference_render_path() is None,
                                     reason="reference render.py neither at /root/reference nor under baseline/_ref")

W_, H_ = 160, 112
BASE_KEYS = {"render", "scaling", "viewspace_points", "visibility_filter", "visible_mask", "selection_mask", "opacity",
             "render_depth", "radii", "render_alphas"}
KEYS_2D = {"render_normals", "render_normals_from_depth", "render_distort"}


def _camera(dev):
    V = scenes.look_at((0.0, -5.5, 3.0), (0.0, 0.0, 0.2)).to(dev)
    K = scenes.intrinsics(W_, H_, 65.0).to(dev)
    return RA.Camera(V, K, W_, H_)


def _anchor_pc(dev, gs_attr):
    tiny = LH.TinyAnchorModel().to(dev)
    tiny.level = tiny.level.to(dev)
    return RA.AnchorModel(tiny, gs_attr=gs_attr)


def _explicit_pc(dev, gs_attr, sh):
    sc, *_ = small_scene(n=4000, sh_degree=sh, width=W_, height=H_, scale=0.1)
    return RA.ExplicitModel(sc.to(dev), gs_attr=gs_attr)


def _run(mod, pc, dev, add_prefilter=True):
    """render() + the backward a training iteration does (train.py:150-200); returns the dict and the gradients"""
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    out = mod.render(_camera(dev), pc, RA.Pipe(add_prefilter), bg)
    g = torch.Generator().manual_seed(41)
    w = torch.rand(3, H_, W_, generator=g).to(dev)
    loss = (out["render"] * w).sum() + out["render_alphas"].sum() + 0.1 * out["render_depth"].sum()
    if pc.gs_attr == "2D":
        wn = torch.rand(1, H_, W_, 3, generator=g).to(dev)
        loss = loss + 0.05 * (out["render_normals"] * wn).sum() + 0.05 * (out["render_normals_from_depth"] * wn[0]).sum()
    loss.backward()
    params = list(pc.m.parameters()) if not pc.explicit_gs else pc.params
    return out, [None if p.grad is None else p.grad.detach().cpu() for p in params]


def _check_contract(out, pc, n_gauss):
    """render.py:96-116: keys, shapes and dtypes of the return dict"""
    assert set(out) == BASE_KEYS | (KEYS_2D if pc.gs_attr == "2D" else set())
    assert out["render"].shape == (3, H_, W_) and out["render_alphas"].shape == (1, H_, W_)
    assert out["render_depth"].shape == (1, H_, W_)
    assert out["radii"].shape == (n_gauss,) and out["radii"].dtype == torch.int32
    assert out["visibility_filter"].dtype == torch.bool and out["visibility_filter"].shape == (n_gauss,)
    assert out["viewspace_points"].shape == (1, n_gauss, 2)
    assert out["viewspace_points"].grad is not None and out["viewspace_points"].grad.shape == (1, n_gauss, 2)
    assert out["scaling"].shape == (n_gauss, 3) and out["opacity"].shape == (n_gauss, 1)
    assert out["visible_mask"].dtype == torch.bool
    if pc.gs_attr == "2D":
        assert out["render_normals"].shape == (1, H_, W_, 3)
        assert out["render_normals_from_depth"].shape[-3:] == (H_, W_, 3)
        assert out["render_distort"].shape == (1, H_, W_, 1)


CASES = [("anchor", "3D", None), ("anchor", "2D", None), ("explicit", "3D", 2), ("explicit", "2D", None)]


def _make_pc(kind, gs_attr, sh, dev):
    return _anchor_pc(dev, gs_attr) if kind == "anchor" else _explicit_pc(dev, gs_attr, sh)


@needs_reference
@pytest.mark.parametrize("kind,gs_attr,sh", CASES)
def test_reference_render_py_runs_on_the_oracle_backend(kind, gs_attr, sh):
    """CPU: the unmodified file drives the oracle; the contract holds and the result equals a direct oracle call"""
    from oracle import gsplat_oracle as O
    mod = RA.load_reference_render("oracle")
    pc = _make_pc(kind, gs_attr, sh, "cpu")
    out, grads = _run(mod, pc, "cpu")
    n = out["radii"].shape[0]
    _check_contract(out, pc, n)
    assert n > 500 and int(out["visibility_filter"].sum()) > 100
    assert all(g is not None and torch.isfinite(g).all() for g in grads)
    # prefilter_voxel (render.py:120-197) alone: anchors whose projection has radius 0 are dropped
    if kind == "anchor":
        cam = _camera("cpu")
        pc.set_anchor_mask(cam.camera_center, cam.resolution_scale)
        vm = mod.prefilter_voxel(cam, pc)
        assert vm.dtype == torch.bool and vm.shape == pc._anchor_mask.shape
        assert torch.equal(vm, out["visible_mask"]) and bool((vm <= pc._anchor_mask).all())
        # without the prefilter render() falls back to the level mask (render.py:27)
        pc2 = _make_pc(kind, gs_attr, sh, "cpu")
        out2, _ = _run(mod, pc2, "cpu", add_prefilter=False)
        assert torch.equal(out2["visible_mask"], pc._anchor_mask)
    # the adapter adds nothing numerically: same image as calling the oracle with the arguments render() builds
    with torch.no_grad():
        cam = _camera("cpu")
        if kind == "anchor":
            xyz, color, opacity, scaling, rot = pc.m.decode(cam.camera_center, out["visible_mask"])
            shd = None
        else:
            xyz, color, opacity, scaling, rot = (p[out["visible_mask"]] for p in pc.params)
            shd = sh
        Kmat = torch.tensor([[cam.fx, 0, cam.cx], [0, cam.fy, cam.cy], [0, 0, 1]], dtype=torch.float32)
        fn = O.rasterization if gs_attr == "3D" else O.rasterization_2dgs
        res = fn(xyz, rot, scaling, opacity.squeeze(-1), color, cam.world_view_transform.T[None], Kmat[None], W_, H_,
                 backgrounds=torch.tensor([[0.1, 0.2, 0.3]]), sh_degree=shd, render_mode="RGB+ED")
        rc = res[0] if gs_attr == "3D" else res[0][0]
        assert torch.equal(out["render"].detach(), rc[0, ..., :3].permute(2, 0, 1))


@needs_reference
@pytest.mark.gpu
@pytest.mark.parametrize("kind,gs_attr,sh", CASES)
def test_reference_render_py_runs_on_the_cuda_drop_in(kind, gs_attr, sh):
    """GPU: the same unmodified file with ``import gsplat`` -> shim/gsplat -> libhgs_raster.so, against its own
    run on the oracle: integer outputs identical, images 1e-4, gradients 1e-3 (north_star tolerances)"""
    ref_mod = RA.load_reference_render("oracle")
    gpu_mod = RA.load_reference_render("shim")
    pc_ref = _make_pc(kind, gs_attr, sh, "cpu")
    pc_gpu = _make_pc(kind, gs_attr, sh, "cuda")
    ro, rg = _run(ref_mod, pc_ref, "cpu")
    go, gg = _run(gpu_mod, pc_gpu, "cuda")
    n = ro["radii"].shape[0]
    _check_contract(go, pc_gpu, n)
    assert torch.equal(go["visible_mask"].cpu(), ro["visible_mask"])
    assert torch.equal(go["radii"].cpu(), ro["radii"]) and torch.equal(go["visibility_filter"].cpu(), ro["visibility_filter"])
    for k in ("render", "render_alphas", "render_depth"):
        err = float(((go[k].detach().cpu() - ro[k].detach()).abs() / ro[k].detach().abs().clamp(min=1.0)).max())
        assert err < 1e-4, (k, err)
    if gs_attr == "2D":
        err = float((go["render_normals"].detach().cpu() - ro["render_normals"].detach()).abs().max())
        assert err < 1e-4, ("render_normals", err)
    for i, (a, b) in enumerate(zip(gg, rg)):
        assert a is not None and b is not None
        assert rel_err(a, b) < 1e-3, (i, rel_err(a, b))
    assert rel_err(go["viewspace_points"].grad.cpu(), ro["viewspace_points"].grad) < 1e-3 or gs_attr == "2D"
