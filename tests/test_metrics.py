"""Metric tail (partial): closed-form checks of e3dge_b200.metrics on CPU."""
import math

import torch

from e3dge_b200 import metrics as M


def test_mse_mae_psnr_closed_form():
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand(2, 3, 256, 256, generator=g) * 2 - 1, torch.rand(2, 3, 256, 256, generator=g) * 2 - 1
    m = M.rec_metrics(a, b)
    mse = ((a - b) ** 2).mean()
    assert torch.allclose(m["loss_l2"], mse) and torch.allclose(m["mae"], (a - b).abs().mean())
    assert abs(m["PSNR"].item() - 10 * math.log10(1.0 / (mse.item() / 4))) < 1e-4  # (x/2+.5): errors halve
    big = torch.rand(1, 3, 1024, 1024, generator=g)
    assert torch.allclose(M.pool_256(big), big.reshape(1, 3, 256, 4, 256, 4).mean((3, 5)), atol=1e-6)
    assert M.pool_256(a) is a


def test_ssim_closed_form_cases():
    x = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    assert abs(M.ssim_loss(x, x).item()) < 1e-6                       # identical images: ssim = 1
    assert torch.allclose(M.ssim_map(x, 1 - x), M.ssim_map(1 - x, x))  # symmetric
    a, b = torch.full((1, 1, 16, 16), 0.3), torch.full((1, 1, 16, 16), 0.7)
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    want = (2 * 0.21 + c1) / (0.09 + 0.49 + c1)  # constant images: variances vanish, the C2 factors cancel
    assert torch.allclose(M.ssim_map(a, b), torch.full((1, 1, 16, 16), want), atol=1e-5)
    y = x.clone().requires_grad_(True)
    M.ssim_loss(y, 1 - x).backward()
    assert torch.isfinite(y.grad).all() and y.grad.abs().max() > 0
