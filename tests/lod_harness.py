"""A miniature of Horizon-GS's LOD anchor model + render() for BASELINE.json configs[3] (test harness only).

It mirrors the *call pattern* of the reference adapter (gaussian_renderer/render.py:16-118 render(),
:120-197 prefilter_voxel()) and of the anchor decode (scene/basic_model.py:297-371
generate_neural_gaussians; scene/lod_model.py:286-290 set_anchor_mask) with a backend switch: the same
Python runs on the CPU oracle and on the CUDA operators.  The decode stays in PyTorch, as in the reference.
"""
import math

import torch
import torch.nn as nn


class TinyAnchorModel(nn.Module):
    def __init__(self, n_anchors=600, n_offsets=10, feat_dim=32, levels=3, extent=3.0, seed=0, voxel0=0.4,
                 standard_dist=8.0, view_dim=3, color_dim=3):
        super().__init__()
        self.view_dim, self.color_dim = view_dim, color_dim      # color_attr 'RGB' (3) or 'SH<d>' (3 (d+1)^2), lod_model.py:58-61
        g = torch.Generator().manual_seed(seed)
        self.n_offsets, self.levels = n_offsets, levels
        self.standard_dist, self.fork = standard_dist, 2
        anchor = (torch.rand(n_anchors, 3, generator=g) * 2 - 1) * extent
        anchor[:, 2] = anchor[:, 2].abs() * 0.3
        self.anchor = nn.Parameter(anchor)
        self.level = torch.randint(0, levels, (n_anchors,), generator=g)
        voxel = voxel0 / (2.0 ** self.level.float())
        self.offset = nn.Parameter(torch.randn(n_anchors, n_offsets, 3, generator=g) * 0.3)
        self.anchor_feat = nn.Parameter(torch.randn(n_anchors, feat_dim, generator=g) * 0.5)
        self.scaling = nn.Parameter(torch.log(voxel)[:, None].expand(-1, 6).contiguous())   # exp() activation
        rot = torch.zeros(n_anchors, 4)
        rot[:, 0] = 1.0
        self.rotation = rot                                                   # identity wxyz (lod_model.py:269-270)
        torch.manual_seed(seed)
        mlp = lambda out: nn.Sequential(nn.Linear(feat_dim + view_dim, feat_dim), nn.ReLU(True), nn.Linear(feat_dim, out))  # noqa: E731
        self.mlp_opacity = nn.Sequential(mlp(n_offsets), nn.Tanh())
        self.mlp_cov = mlp(7 * n_offsets)
        # scene/lod_model.py:80-84 has no activation on the colour MLP; the Sigmoid variant is kept for the RGB harness
        self.mlp_color = nn.Sequential(mlp(3 * n_offsets), nn.Sigmoid()) if color_dim == 3 else mlp(color_dim * n_offsets)

    def anchor_mask(self, cam_center):
        """scene/lod_model.py:286-290 + basic_model.py:192-210: level <= int level of the view distance"""
        dist = (self.anchor.detach() - cam_center).norm(dim=1)
        lvl = torch.log2(self.standard_dist / dist) / math.log2(self.fork)
        int_level = lvl.floor().clamp(0, self.levels - 1).long()
        return self.level.to(self.anchor.device) <= int_level

    def decode(self, cam_center, visible_mask):
        """scene/basic_model.py:297-371 with color_attr == 'RGB', view_dim 3, appearance_dim 0"""
        anchor = self.anchor[visible_mask]
        feat = self.anchor_feat[visible_mask]
        offsets = self.offset[visible_mask]
        scaling = torch.exp(self.scaling[visible_mask])
        view = anchor - cam_center
        view = view / view.norm(dim=1, keepdim=True)
        x = torch.cat([feat, view], 1) if self.view_dim > 0 else feat                # basic_model.py:313-316
        k, cd = self.n_offsets, self.color_dim
        opacity = self.mlp_opacity(x).reshape(-1, 1)
        mask = (opacity > 0).view(-1)
        color = self.mlp_color(x).reshape(-1, cd)
        scale_rot = self.mlp_cov(x).reshape(-1, 7)
        rep = torch.cat([scaling, anchor], -1).repeat_interleave(k, 0)
        allv = torch.cat([rep, color, scale_rot, offsets.reshape(-1, 3)], -1)[mask]
        s_rep, a_rep, color, scale_rot, off = allv.split([6, 3, cd, 7, 3], -1)
        scales = s_rep[:, 3:] * torch.sigmoid(scale_rot[:, :3])
        quats = torch.nn.functional.normalize(scale_rot[:, 3:7])
        xyz = a_rep + off * s_rep[:, :3]
        if cd != 3:
            color = color.reshape(color.shape[0], cd // 3, 3)                       # basic_model.py:368-369
        return xyz, color, opacity[mask], scales, quats


def render(model, viewmat, K, width, height, bg, backend, two_d=False, fused_decode=False):
    """the reference adapter's control flow; `backend` is a module-like object exposing the gsplat names.
    fused_decode: use horizongs_b200.decode.generate_neural_gaussians (csrc/decode.cu) instead of model.decode()."""
    dev = model.anchor.device
    cam_center = torch.linalg.inv(viewmat)[:3, 3]
    if fused_decode:
        # LOD level test + prefilter in one kernel (csrc/project3d.cu anchor_filter_kernel)
        from horizongs_b200 import decode as DEC
        visible = DEC.anchor_visibility(model.anchor, torch.exp(model.scaling.detach()), model.rotation.to(dev), viewmat, K,
                                        int(width), int(height), level=model.level, cam_center=cam_center,
                                        standard_dist=model.standard_dist, fork=model.fork, max_level=model.levels - 1)
    else:
        amask = model.anchor_mask(cam_center)
        # prefilter_voxel(): project the anchors as Gaussians, keep radii > 0 (render.py:120-197)
        means = model.anchor.detach()[amask]
        scales = torch.exp(model.scaling.detach()[amask])[:, :3]
        quats = model.rotation.to(dev)[amask]
        with torch.no_grad():
            if two_d:
                dens = torch.zeros((1, means.shape[0], 2), device=dev)
                proj = backend.fully_fused_projection_2dgs(means, quats, scales, viewmat[None], dens, K[None], int(width),
                                                           int(height), eps2d=0.3, packed=False, near_plane=0.01,
                                                           far_plane=1e10, radius_clip=0.0, sparse_grad=False)
            else:
                proj = backend.fully_fused_projection(means, None, quats, scales, viewmat[None], K[None], int(width),
                                                      int(height), eps2d=0.3, packed=False, near_plane=0.01,
                                                      far_plane=1e10, radius_clip=0.0, sparse_grad=False,
                                                      calc_compensations=False)
        visible = amask.clone()
        visible[amask] = proj[0].squeeze(0) > 0
    if fused_decode:
        from horizongs_b200 import decode as DEC
        xyz, color, opacity, scaling, rot, _ = DEC.generate_neural_gaussians(
            model.anchor, model.anchor_feat, model.offset, torch.exp(model.scaling), cam_center, visible,
            model.mlp_opacity, model.mlp_cov, model.mlp_color)
    else:
        xyz, color, opacity, scaling, rot = model.decode(cam_center, visible)
    # active_sh_degree: None for color_attr 'RGB', else the degree of the SH colours (basic_model.py:371; the harness
    # uses the full degree)
    sh_degree = None if model.color_dim == 3 else math.isqrt(model.color_dim // 3) - 1
    kw = dict(means=xyz, quats=rot, scales=scaling, opacities=opacity.squeeze(-1), colors=color,
              viewmats=viewmat[None], Ks=K[None], backgrounds=bg[None], width=int(width), height=int(height),
              packed=False, sh_degree=sh_degree, render_mode="RGB+ED")
    if two_d:
        (rc, ra, rn, rnd, rd, rm), info = backend.rasterization_2dgs(**kw)
    else:
        rc, ra, info = backend.rasterization(**kw)
    try:                                   # render.py:90-93 (no graph under torch.no_grad())
        info["means2d"].retain_grad()
    except RuntimeError:
        pass
    out = {"render": rc[0, ..., :3].permute(2, 0, 1), "render_depth": rc[0, ..., 3:4].permute(2, 0, 1),
           "render_alphas": ra[0].permute(2, 0, 1), "viewspace_points": info["means2d"],
           "radii": info["radii"].squeeze(0), "visible_mask": visible, "n_gaussians": xyz.shape[0]}
    if two_d:
        out.update(render_normals=rn, render_normals_from_depth=rnd)
    return out
