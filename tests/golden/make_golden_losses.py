"""Generate tests/golden/reference_losses.npz from the REFERENCE's own loss code (utils/loss_utils.py:17-60:
l1_loss, ssim) exactly as train.py:158-160 combines them:
    Ll1 = l1_loss(image, gt_image); ssim_loss = 1.0 - ssim(image, gt_image)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * ssim_loss
Run in the build container (needs /root/reference; it does not exist on the GPU box):
    python tests/golden/make_golden_losses.py
Values and the gradient w.r.t. the rendered image come from the reference code and torch autograd on the CPU
(float64 inputs cast from float32 so that the fixtures are the exact-arithmetic answer for the float32 images)."""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_losses.npz")


def main():
    sys.path.insert(0, REF)
    import utils.loss_utils as lu
    out = {}
    g = torch.Generator().manual_seed(20261018)
    for name, (H, W) in {"a": (37, 53), "b": (16, 16), "c": (9, 70)}.items():
        img = torch.rand(3, H, W, generator=g)
        gt = (img + 0.25 * torch.randn(3, H, W, generator=g)).clamp(0, 1)       # correlated, like a render vs. its target
        x = img.double().requires_grad_()
        y = gt.double()
        l1 = lu.l1_loss(x, y)
        s = lu.ssim(x, y)
        lam = 0.2
        loss = (1.0 - lam) * l1 + lam * (1.0 - s)
        (grad,) = torch.autograd.grad(loss, x)
        out.update({f"{name}_img": img.numpy(), f"{name}_gt": gt.numpy(), f"{name}_l1": np.float64(l1.item()),
                    f"{name}_ssim": np.float64(s.item()), f"{name}_loss": np.float64(loss.item()),
                    f"{name}_grad": grad.numpy(), f"{name}_lambda": np.float64(lam)})
    np.savez(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
