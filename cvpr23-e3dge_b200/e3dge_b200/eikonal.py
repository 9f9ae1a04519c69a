"""Differentiable eikonal term: d sdf / d point WITH a graph, as `get_eikonal_term` builds it with
`autograd.grad(sdf, pts, create_graph=True)` (project/utils/volume_renderer.py:796-802) — the E3DGE encoder
training adds `eikonal_lambda * ((|grad| - 1)^2).mean()` to its loss (trainer.py:618-624, stage1.sh:46-47) and
that loss reaches the encoder through the latents, i.e. through the FiLM table.

Value: the existing kernels (sdf-only forward with stash + e3_siren_points_bwd seeded with 1).
Gradient of a loss L(e) with respect to the FiLM table: with e_bar = dL/de,

    dL/dfilm = d/dfilm [ sum_n e_bar_n . grad_x sdf(x_n) ] = d/dfilm [ D_{e_bar} sdf ]

— the reverse-mode gradient of ONE directional derivative, so no Hessian is formed: a primal + tangent sweep
through the eight FiLM layers (h, dh) and a reverse sweep over both.  Per layer
    arg = gamma (W h + b) + beta,   h' = sin(arg),   dh' = cos(arg) gamma (W dh)
    d/d arg  = a cos(arg) - ad sin(arg) gamma (W dh)          (a, ad = adjoints of h', dh')
    d/d(W dh) = ad cos(arg) gamma
    d gamma  = sum_rows [ d/d arg * (arg - beta) / gamma + ad cos(arg) (W dh) ],   d beta = sum_rows d/d arg.
The 256 x 256 contractions (28 per call, primal and tangent stacked into one operand) run on the tensor cores
through e3_tc_linear_fwd (split-bf16, fp32 accumulate); the element-wise glue between them is device-side
PyTorch.  The generator's weights and the points receive no gradient on this path (frozen generator; the points
come from the camera rays)."""
import torch

from . import _lib


class _PackedLinear:
    """tcgen05 operand images of the seven hidden-layer weights W_l and of their transposes."""

    def __init__(self):
        self.key, self.fwd, self.bwd = None, None, None

    def get(self, net):
        ws = [l.weight for l in net.pts_linears[1:]]
        key = (_lib.pack_epoch,) + tuple((w.data_ptr(), w._version) for w in ws)
        if key == self.key:
            return self.fwd, self.bwd
        lib = _lib.load()
        n = lib.e3_tc_linear_packed_bytes(256, 256) // 4

        def pack(w):
            w = _lib.as_f32c(w.detach())
            buf = torch.empty(n, device=w.device, dtype=torch.float32)
            _lib.check(lib.e3_tc_linear_pack(_lib.ptr(w), 256, 256, _lib.ptr(buf), _lib.cur_stream()),
                       "e3_tc_linear_pack")
            return buf
        self.fwd = [pack(w) for w in ws]
        self.bwd = [pack(w.t()) for w in ws]
        self.key = key
        return self.fwd, self.bwd


def _linear(packed, x):
    """x [M,256] -> x W^T [M,256] on the tensor cores."""
    lib = _lib.load()
    x = _lib.as_f32c(x)
    m = x.shape[0]
    y = torch.empty(m, 256, device=x.device, dtype=torch.float32)
    nbytes = lib.e3_tc_linear_workspace_bytes(m, 256)
    ws = torch.empty(nbytes // 4 + 1, device=x.device, dtype=torch.float32)
    _lib.check(lib.e3_tc_linear_fwd(_lib.ptr(packed), 256, 256, _lib.ptr(x), m, None, _lib.ptr(y), _lib.ptr(ws),
                                    nbytes, _lib.cur_stream()), "e3_tc_linear_fwd")
    return y


class EikonalFn(torch.autograd.Function):
    """(renderer, film [B,9,3,256], points [B,N,3] world space) -> d sdf / d point [B,N,3], differentiable with
    respect to the FiLM table."""

    @staticmethod
    def forward(ctx, renderer, film, points, styles):
        ctx.renderer = renderer
        pts = _lib.as_f32c(points.detach())
        ctx.save_for_backward(film.detach(), pts)
        return renderer.sdf_and_gradient(pts, styles.detach())[1]

    @staticmethod
    @torch.no_grad()
    def backward(ctx, e_bar):
        R = ctx.renderer
        net = R.siren
        film, pts = ctx.saved_tensors
        B, N = pts.shape[0], pts.shape[1]
        M = B * N
        scale = float(R.grid_warper.scale_factor)
        fwd, bwd = net.__dict__.setdefault("_e3_packed_linear", _PackedLinear()).get(net)
        gam = film[:, :8, 0].unsqueeze(2)   # [B,8,1,256]
        bet = film[:, :8, 1].unsqueeze(2)
        betp = film[:, :8, 2].unsqueeze(2)  # gamma * bias + beta
        img = lambda t: t.reshape(B, N, 256)
        x = pts * scale
        t = _lib.as_f32c(e_bar) * scale
        w0 = net.pts_linears[0].weight.detach()            # [256,3]
        args, zds = [], []
        arg = gam[:, 0] * (x @ w0.t()) + betp[:, 0]
        zd = t @ w0.t()
        h, hd = torch.sin(arg), torch.cos(arg) * gam[:, 0] * zd
        args.append(arg), zds.append(zd)
        for l in range(1, 8):                              # primal and tangent share the weights: one GEMM
            z2 = _linear(fwd[l - 1], torch.cat([h.reshape(M, 256), hd.reshape(M, 256)], 0))
            arg = gam[:, l] * img(z2[:M]) + betp[:, l]
            zd = img(z2[M:])
            h, hd = torch.sin(arg), torch.cos(arg) * gam[:, l] * zd
            args.append(arg), zds.append(zd)
        w_sigma = net.sigma_linear.weight.detach().reshape(1, 1, 256) * float(net.sigma_linear.std_init)
        a = torch.zeros(B, N, 256, device=pts.device)
        ad = w_sigma.expand(B, N, 256)
        d_film = torch.zeros_like(film)
        for l in range(7, -1, -1):
            arg, zd, g = args[l], zds[l], gam[:, l]
            c, s = torch.cos(arg), torch.sin(arg)
            p = a * c - ad * s * g * zd
            adc = ad * c
            d_film[:, l, 0] = (p * (arg - bet[:, l]) / g + adc * zd).sum(1)
            d_film[:, l, 1] = p.sum(1)
            if l > 0:
                back = _linear(bwd[l - 1], torch.cat([(g * p).reshape(M, 256), (adc * g).reshape(M, 256)], 0))
                a, ad = img(back[:M]), img(back[M:])
        return None, d_film, None, None


def eikonal_term(renderer, points, styles):
    """d sdf / d point at world-space points [B,N,3]; carries a graph to `styles` when they require grad
    (second-order path), plain values otherwise."""
    if torch.is_grad_enabled() and torch.is_tensor(styles) and styles.requires_grad:
        from .volume_renderer import _FilmFn
        film = _FilmFn.apply(renderer, styles)
        return EikonalFn.apply(renderer, film, points, styles)
    return renderer.sdf_and_gradient(points, styles)[1]
