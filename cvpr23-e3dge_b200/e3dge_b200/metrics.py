"""Image-space metric tail of an inversion frame (partial): what
`LossClass.calc_2d_rec_loss` reports next to its training loss (project/losses/builder.py:130-186) —

    loss_l2 = MSELoss(rgb, gt)                      :142
    mae     = l1_loss(rgb, gt)                      :178
    PSNR    = kornia.metrics.psnr(rgb/2+.5, gt/2+.5, 1.0)        :144-145, 40-41
    SSIM    = 1 - kornia.losses.ssim_loss(rgb, gt, 5)            :169, 180

and the `AdaptiveAvgPool2d((256, 256))` both images go through first when the generator runs at 1024
(datasetgan_runner.py:56-57).  Device-side PyTorch, differentiable.  NOT provided: LPIPS-alex and the ArcFace
identity loss (:143, 150-164: pretrained networks that are not available offline).

Parity status: MSE / MAE / PSNR are closed-form; the SSIM follows kornia's published definition (a
`window_size`-tap Gaussian of sigma 1.5 per axis, reflect padding, C1 = (0.01 max_val)^2, C2 = (0.03 max_val)^2,
loss = mean(clamp((1 - ssim_map) / 2, 0, 1))) but kornia is not installed in the build container, so that one
function is **unpinned** against the reference (tests hold it to closed-form cases only)."""
import math

import torch
import torch.nn.functional as F


def pool_256(images):
    """`pool_256` of the runners: identity at 256^2, exact 4x4 box mean at 1024^2 (AdaptiveAvgPool2d)."""
    return images if images.shape[-2:] == (256, 256) else F.adaptive_avg_pool2d(images, (256, 256))


def psnr(input, target, max_val):
    """kornia.metrics.psnr: 10 log10(max_val^2 / mse), the mse over every element of the batch."""
    return 10.0 * torch.log10(max_val ** 2 / F.mse_loss(input, target, reduction="mean"))


def _gaussian_1d(window_size, sigma, device, dtype):
    x = torch.arange(window_size, device=device, dtype=dtype) - window_size // 2
    if window_size % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2) / (2 * sigma ** 2))
    return g / g.sum()


def ssim_map(img1, img2, window_size=5, max_val=1.0, eps=1e-12):
    """Per-pixel SSIM of two [B,C,H,W] batches (kornia.metrics.ssim: separable Gaussian, sigma 1.5, reflect borders)."""
    c = img1.shape[1]
    k = _gaussian_1d(window_size, 1.5, img1.device, img1.dtype)
    kh, kv = k.view(1, 1, 1, -1).expand(c, 1, 1, -1), k.view(1, 1, -1, 1).expand(c, 1, -1, 1)
    p = window_size // 2

    def blur(t):
        t = F.pad(t, (p, p, p, p), mode="reflect")
        return F.conv2d(F.conv2d(t, kh, groups=c), kv, groups=c)
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mu1, mu2 = blur(img1), blur(img2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1, s2, s12 = blur(img1 * img1) - mu1_sq, blur(img2 * img2) - mu2_sq, blur(img1 * img2) - mu12
    return ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2) + eps)


def ssim_loss(img1, img2, window_size=5, max_val=1.0):
    """kornia.losses.ssim_loss, reduction 'mean'."""
    return torch.clamp((1.0 - ssim_map(img1, img2, window_size, max_val)) / 2, 0, 1).mean()


def rec_metrics(rgb_images, rgb_gt):
    """The metric entries of calc_2d_rec_loss's dict (builder.py:171-181) that need no pretrained network."""
    rgb_images, rgb_gt = pool_256(rgb_images), pool_256(rgb_gt)
    return {"loss_l2": F.mse_loss(rgb_images, rgb_gt), "mae": F.l1_loss(rgb_images, rgb_gt),
            "PSNR": psnr(rgb_images / 2 + 0.5, rgb_gt / 2 + 0.5, 1.0),
            "SSIM": 1 - ssim_loss(rgb_images, rgb_gt, 5)}
