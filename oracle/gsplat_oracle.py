"""Pure-torch CPU restatement of the gsplat rasterization path Horizon-GS calls.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED: gsplat
(third-party, unpinned, environment.yml:28 / README.md:29 of the reference) is
not on this disk; what follows restates the published gsplat ~v1.4 algorithm
("gsplat: An Open-Source Library for Gaussian Splatting", Ye et al. 2024, and
the 2DGS surfel rasterizer of Huang et al. 2024 as integrated in gsplat) and is
anchored on the reference's own call sites:

  gaussian_renderer/render.py:40-54    gsplat.rasterization(...)
  gaussian_renderer/render.py:56-76    gsplat.rasterization_2dgs(...)   (nested return)
  gaussian_renderer/render.py:149-165  fully_fused_projection(means, None, quats, scales, ...)
  gaussian_renderer/render.py:171-186  fully_fused_projection_2dgs(means, quats, scales, viewmats, densifications, ...)

Conventions pinned against in-tree reference code (tests/golden/):
  quaternion wxyz -> rotation       utils/general_utils.py:113-134
  SH basis constants and signs      utils/sh_utils.py:26-112
  world->view matrix                utils/graphics_utils.py:38-49, scene/cameras.py:91

Design notes
  * float32 arithmetic is written as explicit scalar expressions on [N] columns,
    evaluated strictly left to right, never through matmul/einsum, so that a
    CUDA kernel compiled without FMA contraction reproduces every intermediate
    bit for bit (radii are integers derived from floats: a 1-ulp difference
    moves a Gaussian across a tile boundary).
  * Backward passes come from torch autograd; pass float64 tensors to get a
    high-precision gradient reference.
  * Only the un-packed layout ([C, N, ...]) is implemented: the reference always
    passes packed=False (render.py:50,72,158,180).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from . import constants as K


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def _quat_to_rot(quats: torch.Tensor):
    """wxyz quaternion (normalised here) -> rotation entries r[i][j] as [N] columns.

    Same matrix as utils/general_utils.py:113-134 (build_rotation) of the reference.
    """
    w, x, y, z = quats.unbind(-1)
    inv = 1.0 / torch.sqrt(x * x + y * y + z * z + w * w)
    w, x, y, z = w * inv, x * inv, y * inv, z * inv
    x2, y2, z2 = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    return (
        (1.0 - 2.0 * (y2 + z2), 2.0 * (xy - wz), 2.0 * (xz + wy)),
        (2.0 * (xy + wz), 1.0 - 2.0 * (x2 + z2), 2.0 * (yz - wx)),
        (2.0 * (xz - wy), 2.0 * (yz + wx), 1.0 - 2.0 * (x2 + y2)),
    )


def _dot3(a0, b0, a1, b1, a2, b2):
    """(a0*b0 + a1*b1) + a2*b2 -- the one summation order used everywhere."""
    return a0 * b0 + a1 * b1 + a2 * b2


def _scatter_rows(n: int, idx: torch.Tensor, vals: torch.Tensor) -> torch.Tensor:
    """zeros([n, ...]) with rows idx set to vals; differentiable w.r.t. vals."""
    out = torch.zeros((n,) + tuple(vals.shape[1:]), dtype=vals.dtype, device=vals.device)
    return out.index_copy(0, idx, vals)


def _cam_scalars(viewmat: torch.Tensor, Kmat: torch.Tensor):
    R = [[viewmat[i, j] for j in range(3)] for i in range(3)]
    t = [viewmat[i, 3] for i in range(3)]
    fx, fy, cx, cy = Kmat[0, 0], Kmat[1, 1], Kmat[0, 2], Kmat[1, 2]
    return R, t, fx, fy, cx, cy


# --------------------------------------------------------------------------------------
# a3: fully_fused_projection (3DGS)       reference call site render.py:149-165 and inside a5
# --------------------------------------------------------------------------------------
def _project3d_one(means, quats, scales, viewmat, Kmat, width, height, eps2d, near, far, radius_clip):
    N = means.shape[0]
    dt = means.dtype
    R, t, fx, fy, cx, cy = _cam_scalars(viewmat.to(dt), Kmat.to(dt))

    px, py, pz = means.unbind(-1)
    zc_all = _dot3(R[2][0], px, R[2][1], py, R[2][2], pz) + t[2]
    keep = ~((zc_all < near) | (zc_all > far))
    idx = keep.nonzero(as_tuple=True)[0]

    px, py, pz = px[idx], py[idx], pz[idx]
    xc = _dot3(R[0][0], px, R[0][1], py, R[0][2], pz) + t[0]
    yc = _dot3(R[1][0], px, R[1][1], py, R[1][2], pz) + t[1]
    zc = _dot3(R[2][0], px, R[2][1], py, R[2][2], pz) + t[2]

    # covariance  Sigma = (Rq S)(Rq S)^T
    q = _quat_to_rot(quats[idx])
    s = scales[idx].unbind(-1)
    M = [[q[i][j] * s[j] for j in range(3)] for i in range(3)]
    S = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(i, 3):
            S[i][j] = _dot3(M[i][0], M[j][0], M[i][1], M[j][1], M[i][2], M[j][2])
            S[j][i] = S[i][j]
    # camera frame: Sc = Rwc Sigma Rwc^T
    A = [[_dot3(R[i][0], S[0][j], R[i][1], S[1][j], R[i][2], S[2][j]) for j in range(3)] for i in range(3)]
    Sc = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(i, 3):
            Sc[i][j] = _dot3(A[i][0], R[j][0], A[i][1], R[j][1], A[i][2], R[j][2])
            Sc[j][i] = Sc[i][j]

    # perspective projection with the tan-fov clamp on the Jacobian only
    tan_fovx = 0.5 * width / fx
    tan_fovy = 0.5 * height / fy
    lim_x_pos = (width - cx) / fx + K.FOV_MARGIN * tan_fovx
    lim_x_neg = cx / fx + K.FOV_MARGIN * tan_fovx
    lim_y_pos = (height - cy) / fy + K.FOV_MARGIN * tan_fovy
    lim_y_neg = cy / fy + K.FOV_MARGIN * tan_fovy
    rz = 1.0 / zc
    rz2 = rz * rz
    tx = zc * torch.minimum(lim_x_pos, torch.maximum(-lim_x_neg, xc * rz))
    ty = zc * torch.minimum(lim_y_pos, torch.maximum(-lim_y_neg, yc * rz))
    j00 = fx * rz
    j11 = fy * rz
    j02 = -(fx * tx * rz2)
    j12 = -(fy * ty * rz2)
    B00 = j00 * Sc[0][0] + j02 * Sc[2][0]
    B01 = j00 * Sc[0][1] + j02 * Sc[2][1]
    B02 = j00 * Sc[0][2] + j02 * Sc[2][2]
    B11 = j11 * Sc[1][1] + j12 * Sc[2][1]
    B12 = j11 * Sc[1][2] + j12 * Sc[2][2]
    c00 = B00 * j00 + B02 * j02
    c01 = B01 * j11 + B02 * j12
    c11 = B11 * j11 + B12 * j12
    m2x = fx * xc * rz + cx
    m2y = fy * yc * rz + cy

    det_orig = c00 * c11 - c01 * c01
    c00 = c00 + eps2d
    c11 = c11 + eps2d
    det = c00 * c11 - c01 * c01
    with torch.no_grad():
        det_ok = det > 0
    det_safe = torch.where(det_ok, det, torch.ones_like(det))
    comp = torch.sqrt(torch.clamp(det_orig / det_safe, min=0.0))
    inv_det = 1.0 / det_safe
    con_a = c11 * inv_det
    con_b = -c01 * inv_det
    con_c = c00 * inv_det

    with torch.no_grad():
        b = 0.5 * (c00 + c11)
        v1 = b + torch.sqrt(torch.clamp(b * b - det_safe, min=K.EIG_FLOOR))
        radius = torch.ceil(K.RADIUS_SIGMA * torch.sqrt(v1))
        ok = det_ok & ~(radius <= radius_clip)
        ok &= ~((m2x + radius <= 0) | (m2x - radius >= width) | (m2y + radius <= 0) | (m2y - radius >= height))
        radius_i = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)

    def keep_rows(v):
        return torch.where(ok.reshape((-1,) + (1,) * (v.dim() - 1)), v, torch.zeros_like(v))

    radii = _scatter_rows(N, idx, radius_i)
    means2d = _scatter_rows(N, idx, keep_rows(torch.stack([m2x, m2y], -1)))
    depths = _scatter_rows(N, idx, keep_rows(zc))
    conics = _scatter_rows(N, idx, keep_rows(torch.stack([con_a, con_b, con_c], -1)))
    comps = _scatter_rows(N, idx, keep_rows(comp))
    return radii, means2d, depths, conics, comps


def fully_fused_projection(
    means, covars, quats, scales, viewmats, Ks, width, height,
    eps2d=K.EPS2D_DEFAULT, near_plane=K.NEAR_DEFAULT, far_plane=K.FAR_DEFAULT, radius_clip=0.0,
    packed=False, sparse_grad=False, calc_compensations=False,
):
    """-> radii[C,N] i32, means2d[C,N,2], depths[C,N], conics[C,N,3], compensations[C,N]|None.

    Signature as called at render.py:149-165 (covars is the 2nd positional and is None there).
    Culled Gaussians get radii 0 and zeros everywhere else.
    """
    assert covars is None, "the reference always passes quats/scales (render.py:151)"
    assert not packed and not sparse_grad
    outs = [
        _project3d_one(means, quats, scales, viewmats[c], Ks[c], width, height, eps2d, near_plane, far_plane, radius_clip)
        for c in range(viewmats.shape[0])
    ]
    radii, means2d, depths, conics, comps = (torch.stack(x, 0) for x in zip(*outs))
    return radii, means2d, depths, conics, (comps if calc_compensations else None)


# --------------------------------------------------------------------------------------
# a4: fully_fused_projection_2dgs (surfels)   reference call site render.py:171-186 and inside a6
# --------------------------------------------------------------------------------------
def _project2d_one(means, quats, scales, viewmat, Kmat, width, height, near, far, radius_clip):
    N = means.shape[0]
    dt = means.dtype
    R, t, fx, fy, cx, cy = _cam_scalars(viewmat.to(dt), Kmat.to(dt))
    px, py, pz = means.unbind(-1)
    zc_all = _dot3(R[2][0], px, R[2][1], py, R[2][2], pz) + t[2]
    keep = ~((zc_all < near) | (zc_all > far))
    idx = keep.nonzero(as_tuple=True)[0]
    px, py, pz = px[idx], py[idx], pz[idx]
    mc = [_dot3(R[i][0], px, R[i][1], py, R[i][2], pz) + t[i] for i in range(3)]

    q = _quat_to_rot(quats[idx])
    s = scales[idx].unbind(-1)
    RQ = [[_dot3(R[i][0], q[0][j], R[i][1], q[1][j], R[i][2], q[2][j]) for j in range(3)] for i in range(3)]
    # WH columns: s0 * tangent-u, s1 * tangent-v, centre (third scale ignored)
    WH = [[RQ[i][0] * s[0], RQ[i][1] * s[1], mc[i]] for i in range(3)]
    M0 = [fx * WH[0][j] + cx * WH[2][j] for j in range(3)]
    M1 = [fy * WH[1][j] + cy * WH[2][j] for j in range(3)]
    M2 = [WH[2][j] for j in range(3)]

    dist = M2[0] * M2[0] + M2[1] * M2[1] - M2[2] * M2[2]
    with torch.no_grad():
        dist_ok = dist != 0
    dist_safe = torch.where(dist_ok, dist, torch.ones_like(dist))
    invd = 1.0 / dist_safe
    f = (invd, invd, -invd)
    m2x = f[0] * M0[0] * M2[0] + f[1] * M0[1] * M2[1] + f[2] * M0[2] * M2[2]
    m2y = f[0] * M1[0] * M2[0] + f[1] * M1[1] * M2[1] + f[2] * M1[2] * M2[2]
    with torch.no_grad():
        tmpx = f[0] * M0[0] * M0[0] + f[1] * M0[1] * M0[1] + f[2] * M0[2] * M0[2]
        tmpy = f[0] * M1[0] * M1[0] + f[1] * M1[1] * M1[1] + f[2] * M1[2] * M1[2]
        hx = m2x * m2x - tmpx
        hy = m2y * m2y - tmpy
        radius = torch.ceil(K.RADIUS_SIGMA * torch.sqrt(torch.clamp(torch.maximum(hx, hy), min=K.RADIUS_FLOOR_2DGS)))
        ok = dist_ok & ~(radius <= radius_clip)
        ok &= ~((m2x + radius <= 0) | (m2x - radius >= width) | (m2y + radius <= 0) | (m2y - radius >= height))
        radius_i = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)
        # normal = third rotation column in camera frame, flipped to face the camera
        flip = _dot3(-RQ[0][2], mc[0], -RQ[1][2], mc[1], -RQ[2][2], mc[2]) > 0
        sign = torch.where(flip, torch.ones_like(dist), -torch.ones_like(dist))

    def keep_rows(v):
        return torch.where(ok.reshape((-1,) + (1,) * (v.dim() - 1)), v, torch.zeros_like(v))

    normal = torch.stack([RQ[0][2] * sign, RQ[1][2] * sign, RQ[2][2] * sign], -1)
    rt = torch.stack([torch.stack(M0, -1), torch.stack(M1, -1), torch.stack(M2, -1)], -2)  # [n,3,3]
    radii = _scatter_rows(N, idx, radius_i)
    means2d = _scatter_rows(N, idx, keep_rows(torch.stack([m2x, m2y], -1)))
    depths = _scatter_rows(N, idx, keep_rows(mc[2]))
    ray_transforms = _scatter_rows(N, idx, keep_rows(rt))
    normals = _scatter_rows(N, idx, keep_rows(normal))
    return radii, means2d, depths, ray_transforms, normals


def fully_fused_projection_2dgs(
    means, quats, scales, viewmats, densifications, Ks, width, height,
    eps2d=K.EPS2D_DEFAULT, near_plane=K.NEAR_DEFAULT, far_plane=K.FAR_DEFAULT, radius_clip=0.0,
    packed=False, sparse_grad=False,
):
    """-> radii[C,N] i32, means2d[C,N,2], depths[C,N], ray_transforms[C,N,3,3], normals[C,N,3].

    Positional order as called at render.py:171-186 (fork-specific ``densifications``
    argument in 5th position; it only carries a gradient slot and is not read).
    """
    assert not packed and not sparse_grad
    outs = [
        _project2d_one(means, quats, scales, viewmats[c], Ks[c], width, height, near_plane, far_plane, radius_clip)
        for c in range(viewmats.shape[0])
    ]
    return tuple(torch.stack(x, 0) for x in zip(*outs))


# --------------------------------------------------------------------------------------
# a7: spherical harmonics                  conventions: utils/sh_utils.py:57-112
# --------------------------------------------------------------------------------------
def _sh_bases(degree: int, dirs: torch.Tensor):
    """Real SH basis values b_k(dir) for k < (degree+1)^2, as a list of [...] tensors.

    dirs are normalised here.  Polynomial forms follow Sloan, "Efficient Spherical
    Harmonic Evaluation" (JCGT 2013), which is what gsplat evaluates; they equal
    the explicit forms in utils/sh_utils.py:74-112 for unit vectors.
    """
    x, y, z = dirs.unbind(-1)
    inorm = 1.0 / torch.sqrt(x * x + y * y + z * z)
    x, y, z = x * inorm, y * inorm, z * inorm
    b = [torch.full_like(x, 0.2820947917738781)]
    if degree < 1:
        return b
    b += [-0.48860251190292 * y, 0.48860251190292 * z, -0.48860251190292 * x]
    if degree < 2:
        return b
    z2 = z * z
    fTmp0B = -1.092548430592079 * z
    fC1 = x * x - y * y
    fS1 = 2.0 * x * y
    pSH6 = 0.9461746957575601 * z2 - 0.3153915652525201
    b += [0.5462742152960395 * fS1, fTmp0B * y, pSH6, fTmp0B * x, 0.5462742152960395 * fC1]
    if degree < 3:
        return b
    fTmp0C = -2.285228997322329 * z2 + 0.4570457994644658
    fTmp1B = 1.445305721320277 * z
    fC2 = x * fC1 - y * fS1
    fS2 = x * fS1 + y * fC1
    pSH12 = z * (1.865881662950577 * z2 - 1.119528997770346)
    b += [-0.5900435899266435 * fS2, fTmp1B * fS1, fTmp0C * y, pSH12, fTmp0C * x, fTmp1B * fC1,
          -0.5900435899266435 * fC2]
    if degree < 4:
        return b
    fTmp0D = z * (-4.683325804901025 * z2 + 2.007139630671868)
    fTmp1C = 3.31161143515146 * z2 - 0.47308734787878
    fTmp2B = -1.770130769779931 * z
    fC3 = x * fC2 - y * fS2
    fS3 = x * fS2 + y * fC2
    pSH20 = 1.984313483298443 * z * pSH12 + -1.006230589874905 * pSH6
    b += [0.6258357354491763 * fS3, fTmp2B * fS2, fTmp1C * fS1, fTmp0D * y, pSH20, fTmp0D * x,
          fTmp1C * fC1, fTmp2B * fC2, 0.6258357354491763 * fC3]
    return b


def spherical_harmonics(degrees_to_use: int, dirs: torch.Tensor, coeffs: torch.Tensor,
                        masks: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dirs[...,3] (un-normalised), coeffs[...,K,3] -> colors[...,3]; zeros where ~masks.

    coeffs layout is [N, K, 3] as produced at scene/basic_model.py:369,378.
    """
    assert (degrees_to_use + 1) ** 2 <= coeffs.shape[-2]
    bases = _sh_bases(degrees_to_use, dirs)
    out = bases[0][..., None] * coeffs[..., 0, :]
    for k in range(1, len(bases)):
        out = out + bases[k][..., None] * coeffs[..., k, :]
    if masks is not None:
        out = torch.where(masks[..., None], out, torch.zeros_like(out))
    return out


# --------------------------------------------------------------------------------------
# a8-a10: tile intersection, 64-bit key sort, per-tile ranges   (integer, bit-exact stage)
# --------------------------------------------------------------------------------------
def _n_bits(n: int) -> int:
    return int(math.floor(math.log2(n))) + 1


def isect_tiles(means2d, radii, depths, tile_size, tile_width, tile_height, sort=True):
    """-> tiles_per_gauss[C,N] i32, isect_ids[I] i64, flatten_ids[I] i32.

    key = cam_id << (32 + tile_bits) | tile_id << 32 | int32 bits of depth;
    value = flat index c*N + n; tiles visited row-major; stable ascending sort.
    """
    C, N = radii.shape
    f32 = torch.float32
    m = means2d.detach().to(f32).reshape(C * N, 2)
    r = radii.reshape(C * N)
    d = depths.detach().to(f32).reshape(C * N)
    ts = float(tile_size)
    tile_r = r.to(f32) / ts
    tx, ty = m[:, 0] / ts, m[:, 1] / ts
    xmin = torch.floor(tx - tile_r).clamp(0, tile_width).to(torch.int64)
    ymin = torch.floor(ty - tile_r).clamp(0, tile_height).to(torch.int64)
    xmax = torch.ceil(tx + tile_r).clamp(0, tile_width).to(torch.int64)
    ymax = torch.ceil(ty + tile_r).clamp(0, tile_height).to(torch.int64)
    counts = (ymax - ymin) * (xmax - xmin)
    counts = torch.where(r > 0, counts, torch.zeros_like(counts))
    tiles_per_gauss = counts.to(torch.int32).reshape(C, N)

    n_tiles = tile_width * tile_height
    tile_bits = _n_bits(n_tiles)
    total = int(counts.sum())
    flat = torch.repeat_interleave(torch.arange(C * N), counts)
    start = torch.cumsum(counts, 0) - counts
    local = torch.arange(total) - start[flat]
    w = (xmax - xmin)[flat]
    tile_id = (ymin[flat] + local // w.clamp(min=1)) * tile_width + (xmin[flat] + local % w.clamp(min=1))
    cam = flat // N
    depth_bits = d.view(torch.int32).to(torch.int64)[flat]  # sign-extending cast, as gsplat does
    isect_ids = (cam << (32 + tile_bits)) | (tile_id << 32) | depth_bits
    flatten_ids = flat.to(torch.int32)
    if sort:
        isect_ids, order = torch.sort(isect_ids, stable=True)
        flatten_ids = flatten_ids[order]
    return tiles_per_gauss, isect_ids, flatten_ids


def isect_offset_encode(isect_ids, n_cameras, tile_width, tile_height):
    """-> offsets[C, tile_height, tile_width] i32: first sorted index of each (cam, tile);
    empty tiles inherit the next start, trailing ones get I."""
    n_tiles = tile_width * tile_height
    tile_bits = _n_bits(n_tiles)
    hi = isect_ids >> 32
    flat = (hi >> tile_bits) * n_tiles + (hi & ((1 << tile_bits) - 1))
    off = torch.searchsorted(flat.contiguous(), torch.arange(n_cameras * n_tiles), right=False)
    return off.to(torch.int32).reshape(n_cameras, tile_height, tile_width)


# --------------------------------------------------------------------------------------
# a11 / a12: alpha blending
# --------------------------------------------------------------------------------------
def _blend_chunks(alpha_fn, G, feats, P, dt, chunk, extra=None):
    """Front-to-back blend of G sorted Gaussians over P pixels, in chunks, with the
    kernel's exact termination rule.  alpha_fn(s, e) -> (alpha[P,g], valid[P,g]).
    extra(T_before, vis, included, s, e) lets the 2DGS path accumulate more outputs.
    Returns acc[P,D], T[P].
    """
    T = torch.ones(P, dtype=dt)
    done = torch.zeros(P, dtype=torch.bool)
    acc = torch.zeros(P, feats.shape[-1], dtype=dt)
    for s in range(0, G, chunk):
        e = min(G, s + chunk)
        alpha, valid = alpha_fn(s, e)
        a_eff = torch.where(valid, alpha, torch.zeros_like(alpha))
        # sequential product starting from the carried T: ((T*(1-a0))*(1-a1))*...
        Tincl = torch.cumprod(torch.cat([T[:, None], 1.0 - a_eff], 1), 1)
        Tbefore, Tincl = Tincl[:, :-1], Tincl[:, 1:]
        with torch.no_grad():
            term_here = valid & (Tincl <= K.T_EPS)
            terminated = torch.cummax(term_here.to(torch.uint8), 1).values.bool()
            included = valid & ~terminated & ~done[:, None]
            any_term = term_here.any(1)
            first = term_here.to(torch.uint8).argmax(1, keepdim=True)
        vis = torch.where(included, a_eff * Tbefore, torch.zeros_like(alpha))
        acc = acc + vis @ feats[s:e]
        if extra is not None:
            extra(Tbefore, vis, included, s, e)
        T_new = torch.where(any_term, Tbefore.gather(1, first)[:, 0], Tincl[:, -1])
        T = torch.where(done, T, T_new)
        done = done | any_term
        if bool(done.all()):
            break
    return acc, T


def _tile_pixels(ty, tx, tile_size, dt):
    ii = torch.arange(tile_size, dtype=dt)
    py = (ty * tile_size + ii + 0.5)[:, None].expand(tile_size, tile_size).reshape(-1)
    px = (tx * tile_size + ii + 0.5)[None, :].expand(tile_size, tile_size).reshape(-1)
    return px, py


def _assemble(tiles, th, tw, ts, H, W):
    """list (row-major over tiles) of [ts*ts, D] -> [H, W, D]"""
    D = tiles[0].shape[-1]
    img = torch.stack(tiles, 0).reshape(th, tw, ts, ts, D).permute(0, 2, 1, 3, 4).reshape(th * ts, tw * ts, D)
    return img[:H, :W]


def rasterize_to_pixels(means2d, conics, colors, opacities, image_width, image_height, tile_size,
                        isect_offsets, flatten_ids, backgrounds=None, tile_subset=None, chunk=256):
    """-> render_colors[C,H,W,D], render_alphas[C,H,W,1].

    sigma = 0.5*(a dx^2 + c dy^2) + b dx dy at pixel centres (x+0.5, y+0.5);
    alpha = min(0.999, o * exp(-sigma)); skipped when sigma < 0 or alpha < 1/255;
    a pixel stops *before* the Gaussian that would bring T to <= 1e-4.
    tile_subset: optional iterable of (cam, ty, tx) to evaluate (others stay 0) --
    used for the bounded CPU-baseline sample.
    """
    C, N = opacities.shape
    dt = means2d.dtype
    D = colors.shape[-1]
    th, tw = isect_offsets.shape[1:]
    n_isects = flatten_ids.shape[0]
    offs = isect_offsets.reshape(-1).tolist() + [n_isects]
    m2 = means2d.reshape(C * N, 2)
    cn = conics.reshape(C * N, 3)
    op = opacities.reshape(C * N)
    col = colors.reshape(C * N, D)
    wanted = None if tile_subset is None else set(tile_subset)
    imgs, alphas = [], []
    for c in range(C):
        ctiles, atiles = [], []
        for ty in range(th):
            for tx in range(tw):
                t = (c * th + ty) * tw + tx
                s0, e0 = offs[t], offs[t + 1]
                if e0 <= s0 or (wanted is not None and (c, ty, tx) not in wanted):
                    ctiles.append(torch.zeros(tile_size * tile_size, D, dtype=dt))
                    atiles.append(torch.ones(tile_size * tile_size, 1, dtype=dt))
                    continue
                g = flatten_ids[s0:e0].long()
                px, py = _tile_pixels(ty, tx, tile_size, dt)
                xy, cc, oo = m2[g], cn[g], op[g]

                def alpha_fn(s, e):
                    dx = xy[s:e, 0][None, :] - px[:, None]
                    dy = xy[s:e, 1][None, :] - py[:, None]
                    sigma = 0.5 * (cc[s:e, 0] * dx * dx + cc[s:e, 2] * dy * dy) + cc[s:e, 1] * dx * dy
                    alpha = torch.clamp(oo[s:e] * torch.exp(-sigma), max=K.ALPHA_MAX)
                    with torch.no_grad():
                        valid = (sigma >= 0) & (alpha >= K.ALPHA_MIN)
                    return alpha, valid

                acc, T = _blend_chunks(alpha_fn, e0 - s0, col[g], tile_size * tile_size, dt, chunk)
                ctiles.append(acc)
                atiles.append(T[:, None])
        img = _assemble(ctiles, th, tw, tile_size, image_height, image_width)
        Timg = _assemble(atiles, th, tw, tile_size, image_height, image_width)
        if backgrounds is not None:
            img = img + Timg * backgrounds[c].to(dt)
        imgs.append(img)
        alphas.append(1.0 - Timg)
    return torch.stack(imgs, 0), torch.stack(alphas, 0)


def rasterize_to_pixels_2dgs(means2d, ray_transforms, colors, opacities, normals, image_width, image_height,
                             tile_size, isect_offsets, flatten_ids, backgrounds=None, distloss=False,
                             tile_subset=None, chunk=256):
    """-> render_colors[C,H,W,D], render_alphas[C,H,W,1], render_normals[C,H,W,3],
          render_distort[C,H,W,1] (zeros unless distloss), render_median[C,H,W,1].

    Per pixel (x,y): h_u = x*M2 - M0, h_v = y*M2 - M1, p = h_u x h_v, s = p.xy/p.z;
    weight = min(|s|^2, 2*|mean2d - px|^2); sigma = weight/2; same alpha rules as 3DGS.
    The last colour channel is depth for the distortion / median outputs.
    """
    C, N = opacities.shape
    dt = means2d.dtype
    D = colors.shape[-1]
    th, tw = isect_offsets.shape[1:]
    n_isects = flatten_ids.shape[0]
    offs = isect_offsets.reshape(-1).tolist() + [n_isects]
    m2 = means2d.reshape(C * N, 2)
    rt = ray_transforms.reshape(C * N, 3, 3)
    op = opacities.reshape(C * N)
    col = colors.reshape(C * N, D)
    nor = normals.reshape(C * N, 3)
    wanted = None if tile_subset is None else set(tile_subset)
    P = tile_size * tile_size
    outs = [[] for _ in range(5)]
    for c in range(C):
        tl = [[] for _ in range(5)]
        for ty in range(th):
            for tx in range(tw):
                t = (c * th + ty) * tw + tx
                s0, e0 = offs[t], offs[t + 1]
                if e0 <= s0 or (wanted is not None and (c, ty, tx) not in wanted):
                    for k, v in enumerate((torch.zeros(P, D), torch.ones(P, 1), torch.zeros(P, 3),
                                           torch.zeros(P, 1), torch.zeros(P, 1))):
                        tl[k].append(v.to(dt))
                    continue
                g = flatten_ids[s0:e0].long()
                px, py = _tile_pixels(ty, tx, tile_size, dt)
                xy, Ms, oo = m2[g], rt[g], op[g]
                feats = torch.cat([col[g], nor[g]], -1)
                depth_g = col[g][:, -1]
                state = {
                    "distort": torch.zeros(P, dtype=dt),
                    "acc_vd": torch.zeros(P, dtype=dt),
                    "acc_w": torch.zeros(P, dtype=dt),
                    "median": torch.zeros(P, dtype=dt),
                }

                def alpha_fn(s, e):
                    uM, vM, wM = Ms[s:e, 0], Ms[s:e, 1], Ms[s:e, 2]          # [g,3]
                    hu = px[:, None, None] * wM[None] - uM[None]              # [P,g,3]
                    hv = py[:, None, None] * wM[None] - vM[None]
                    cx_ = hu[..., 1] * hv[..., 2] - hu[..., 2] * hv[..., 1]
                    cy_ = hu[..., 2] * hv[..., 0] - hu[..., 0] * hv[..., 2]
                    cz_ = hu[..., 0] * hv[..., 1] - hu[..., 1] * hv[..., 0]
                    with torch.no_grad():
                        z_ok = cz_ != 0
                    cz_s = torch.where(z_ok, cz_, torch.ones_like(cz_))
                    sx, sy = cx_ / cz_s, cy_ / cz_s
                    w3 = sx * sx + sy * sy
                    dx = xy[s:e, 0][None, :] - px[:, None]
                    dy = xy[s:e, 1][None, :] - py[:, None]
                    w2 = K.FILTER_INV_SQUARE_2DGS * (dx * dx + dy * dy)
                    sigma = 0.5 * torch.minimum(w3, w2)
                    alpha = torch.clamp(oo[s:e] * torch.exp(-sigma), max=K.ALPHA_MAX)
                    with torch.no_grad():
                        valid = z_ok & (sigma >= 0) & (alpha >= K.ALPHA_MIN)
                    return alpha, valid

                def extra(Tbefore, vis, included, s, e):
                    dg = depth_g[s:e][None, :]
                    if distloss:
                        # sequential recurrences of the kernel, closed form per chunk:
                        #   distort += 2*(vis*d*(1-T) - vis*acc_vd);  acc_vd += vis*d
                        vd = vis * dg
                        cum_vd = torch.cumsum(vd, 1) - vd + state["acc_vd"][:, None]
                        state["distort"] = state["distort"] + (2.0 * (vd * (1.0 - Tbefore) - vis * cum_vd)).sum(1)
                        state["acc_vd"] = state["acc_vd"] + vd.sum(1)
                    with torch.no_grad():
                        hit = included & (Tbefore > K.MEDIAN_T_2DGS)
                        anyhit = hit.any(1)
                        last = (hit.shape[1] - 1) - hit.flip(1).to(torch.uint8).argmax(1)
                    med = dg.expand_as(vis).gather(1, last[:, None])[:, 0]
                    state["median"] = torch.where(anyhit, med, state["median"])

                acc, T = _blend_chunks(alpha_fn, e0 - s0, feats, P, dt, chunk, extra)
                tl[0].append(acc[:, :D])
                tl[1].append(T[:, None])
                tl[2].append(acc[:, D:])
                tl[3].append(state["distort"][:, None])
                tl[4].append(state["median"][:, None])
        img, Timg, nimg, dimg, mimg = (_assemble(x, th, tw, tile_size, image_height, image_width) for x in tl)
        if backgrounds is not None:
            img = img + Timg * backgrounds[c].to(dt)
        for k, v in enumerate((img, 1.0 - Timg, nimg, dimg, mimg)):
            outs[k].append(v)
    return tuple(torch.stack(x, 0) for x in outs)


# --------------------------------------------------------------------------------------
# a13: post-ops
# --------------------------------------------------------------------------------------
def depth_to_normal(depths, camtoworlds, Ks):
    """depths[C,H,W,1] (z-depth) -> finite-difference world-space normals[C,H,W,3], border = 0."""
    C, H, W, _ = depths.shape
    dt = depths.dtype
    x, y = torch.meshgrid(torch.arange(W, dtype=dt), torch.arange(H, dtype=dt), indexing="xy")
    out = []
    for c in range(C):
        fx, fy, cx, cy = Ks[c, 0, 0], Ks[c, 1, 1], Ks[c, 0, 2], Ks[c, 1, 2]
        dirs_c = torch.stack([(x - cx + 0.5) / fx, (y - cy + 0.5) / fy, torch.ones_like(x)], -1)
        dirs_w = dirs_c @ camtoworlds[c, :3, :3].to(dt).T
        pts = camtoworlds[c, :3, 3].to(dt) + depths[c] * dirs_w
        dx = pts[2:, 1:-1] - pts[:-2, 1:-1]
        dy = pts[1:-1, 2:] - pts[1:-1, :-2]
        n = F.normalize(torch.cross(dx, dy, dim=-1), dim=-1)
        out.append(F.pad(n, (0, 0, 1, 1, 1, 1), value=0.0))
    return torch.stack(out, 0)


# --------------------------------------------------------------------------------------
# a5 / a6: the two public pipelines        reference call sites render.py:40-54 and :56-76
# --------------------------------------------------------------------------------------
def _view_colors(means, colors, viewmats, radii, sh_degree):
    C = viewmats.shape[0]
    if sh_degree is None:
        return colors[None].expand(C, -1, -1) if colors.dim() == 2 else colors
    c2w = torch.linalg.inv(viewmats.to(means.dtype))
    dirs = means[None, :, :] - c2w[:, None, :3, 3]
    shs = colors[None].expand(C, -1, -1, -1) if colors.dim() == 3 else colors
    out = spherical_harmonics(sh_degree, dirs, shs, masks=radii > 0)
    return torch.clamp_min(out + K.SH_OFFSET, 0.0)


def _with_depth_channel(colors, depths, backgrounds, render_mode):
    if render_mode in ("RGB+D", "RGB+ED"):
        colors = torch.cat([colors, depths[..., None]], -1)
        if backgrounds is not None:
            backgrounds = torch.cat([backgrounds, torch.zeros_like(backgrounds[:, :1])], -1)
    elif render_mode in ("D", "ED"):
        colors = depths[..., None]
        if backgrounds is not None:
            backgrounds = torch.zeros_like(backgrounds[:, :1])
    return colors, backgrounds


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                  near_plane=K.NEAR_DEFAULT, far_plane=K.FAR_DEFAULT, radius_clip=0.0, eps2d=K.EPS2D_DEFAULT,
                  sh_degree=None, packed=False, tile_size=K.TILE_SIZE, backgrounds=None, render_mode="RGB",
                  sparse_grad=False, absgrad=False, rasterize_mode="classic", tile_subset=None):
    """3DGS pipeline, signature and outputs as used at render.py:40-54.
    -> render_colors[C,H,W,3|4|1], render_alphas[C,H,W,1], meta."""
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    assert not packed and not sparse_grad and not absgrad
    C = viewmats.shape[0]
    radii, means2d, depths, conics, comps = fully_fused_projection(
        means, None, quats, scales, viewmats, Ks, width, height, eps2d=eps2d, near_plane=near_plane,
        far_plane=far_plane, radius_clip=radius_clip, calc_compensations=(rasterize_mode == "antialiased"))
    opac = opacities[None].expand(C, -1)
    if comps is not None:
        opac = opac * comps
    tw, th = math.ceil(width / tile_size), math.ceil(height / tile_size)
    tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tw, th)
    isect_offsets = isect_offset_encode(isect_ids, C, tw, th)
    feats = _view_colors(means, colors, viewmats, radii, sh_degree)
    feats, bgs = _with_depth_channel(feats, depths, backgrounds, render_mode)
    render_colors, render_alphas = rasterize_to_pixels(
        means2d, conics, feats, opac, width, height, tile_size, isect_offsets, flatten_ids,
        backgrounds=bgs, tile_subset=tile_subset)
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=K.ED_ALPHA_FLOOR)], -1)
    meta = dict(camera_ids=None, gaussian_ids=None, radii=radii, means2d=means2d, depths=depths, conics=conics,
                opacities=opac, tile_width=tw, tile_height=th, tiles_per_gauss=tiles_per_gauss,
                isect_ids=isect_ids, flatten_ids=flatten_ids, isect_offsets=isect_offsets,
                width=width, height=height, tile_size=tile_size, n_cameras=C)
    return render_colors, render_alphas, meta


def rasterization_2dgs(means, quats, scales, opacities, colors, viewmats, Ks, width, height,
                       near_plane=K.NEAR_DEFAULT, far_plane=K.FAR_DEFAULT, radius_clip=0.0, eps2d=K.EPS2D_DEFAULT,
                       sh_degree=None, packed=False, tile_size=K.TILE_SIZE, backgrounds=None, render_mode="RGB",
                       sparse_grad=False, absgrad=False, distloss=False, depth_mode="expected", tile_subset=None):
    """2DGS pipeline; returns the NESTED form the reference unpacks at render.py:56-76:
    ((colors, alphas, normals, normals_from_depth, distort, median), meta)."""
    assert render_mode in ("RGB", "D", "ED", "RGB+D", "RGB+ED")
    assert not packed and not sparse_grad and not absgrad
    C = viewmats.shape[0]
    radii, means2d, depths, ray_transforms, normals = fully_fused_projection_2dgs(
        means, quats, scales, viewmats, None, Ks, width, height, eps2d=eps2d, near_plane=near_plane,
        far_plane=far_plane, radius_clip=radius_clip)
    opac = opacities[None].expand(C, -1)
    tw, th = math.ceil(width / tile_size), math.ceil(height / tile_size)
    tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(means2d, radii, depths, tile_size, tw, th)
    isect_offsets = isect_offset_encode(isect_ids, C, tw, th)
    feats = _view_colors(means, colors, viewmats, radii, sh_degree)
    feats, bgs = _with_depth_channel(feats, depths, backgrounds, render_mode)
    render_colors, render_alphas, render_normals, render_distort, render_median = rasterize_to_pixels_2dgs(
        means2d, ray_transforms, feats, opac, normals, width, height, tile_size, isect_offsets, flatten_ids,
        backgrounds=bgs, distloss=distloss, tile_subset=tile_subset)
    render_normals_from_depth = None
    if render_mode in ("ED", "RGB+ED"):
        render_colors = torch.cat(
            [render_colors[..., :-1], render_colors[..., -1:] / render_alphas.clamp(min=K.ED_ALPHA_FLOOR)], -1)
    c2w = torch.linalg.inv(viewmats.to(means.dtype))
    if render_mode in ("RGB+D", "RGB+ED"):
        d4n = render_colors[..., -1:] if depth_mode == "expected" else render_median
        render_normals_from_depth = depth_to_normal(d4n, c2w, Ks.to(means.dtype)).squeeze(0)
    render_normals = torch.einsum("cij,chwj->chwi", c2w[:, :3, :3], render_normals)
    meta = dict(camera_ids=None, gaussian_ids=None, radii=radii, means2d=means2d, depths=depths,
                ray_transforms=ray_transforms, normals=normals, opacities=opac, tile_width=tw, tile_height=th,
                tiles_per_gauss=tiles_per_gauss, isect_ids=isect_ids, flatten_ids=flatten_ids,
                isect_offsets=isect_offsets, width=width, height=height, tile_size=tile_size, n_cameras=C,
                render_distort=render_distort)
    return (render_colors, render_alphas, render_normals, render_normals_from_depth, render_distort,
            render_median), meta
