"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch, fp32 / fp64) of the per-sample MLP tail of the E3DGE
local branch: SFT fusion of the 2-D-aligned and 3-D-projected features, positional encoding of the sample
position, and the ResnetBlockFC that maps the 301-d local feature to the renderer's texture modulation.

Follows, line by line:
  * `ResnetBlockFC.forward`          project/models/helper_modules/resnetfc.py:53-62
  * `Fuse_sft_MLP.forward`           project/models/helper_modules/sft.py:103-109
  * `PosEncoding.forward`            project/utils/misc_utils.py:166-184   (N_freqs = 7, log scale)
  * call site                        project/trainers/E3DGE/e3dge_full_runner.py:282-297
  * alpha / beta split               project/utils/volume_renderer.py:327-336

Pinned by tests/golden/local_mlp.npz, recorded from the reference's own modules by
oracle/gen_golden_local_mlp.py (tests/test_oracle_golden.py::test_local_mlp_case)."""
import torch
import torch.nn.functional as F

N_FREQS = 7
FUSE = "fuse_sft_block."
TEX = "renderer.network.netLocal.local_feat_to_tex_modulations_linear."


def resnet_block_fc(x, sd, key):
    """x_s + fc_1(relu(fc_0(relu(x)))), x_s = shortcut(x) (no bias) — resnetfc.py:53-62."""
    net = F.linear(F.relu(x), sd[key + "fc_0.weight"], sd[key + "fc_0.bias"])
    dx = F.linear(F.relu(net), sd[key + "fc_1.weight"], sd[key + "fc_1.bias"])
    x_s = F.linear(x, sd[key + "shortcut.weight"]) if key + "shortcut.weight" in sd else x
    return x_s + dx


def fuse_sft_mlp(enc_feat, dec_feat, sd, key=FUSE, w=1):
    """dec + w * (dec * scale(e) + shift(e)), e = ResnetBlockFC(cat[enc, dec]) — sft.py:103-109."""
    e = resnet_block_fc(torch.cat([enc_feat, dec_feat], -1), sd, key + "encode_enc.")

    def branch(name):
        h = F.leaky_relu(F.linear(e, sd[f"{key}{name}.0.weight"], sd[f"{key}{name}.0.bias"]), 0.2)
        return F.linear(h, sd[f"{key}{name}.2.weight"], sd[f"{key}{name}.2.bias"])
    return dec_feat + w * (dec_feat * branch("scale") + branch("shift"))


def pos_encoding(x, n_freqs=N_FREQS):
    """(x, sin(2^k x), cos(2^k x))_k — misc_utils.py:166-184."""
    out = [x]
    for k in range(n_freqs):
        f = 2.0 ** k
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


def local_feats(feat_2d, feat_3d, points, sd):
    """The 301-d `feats` the runner hands to the renderer (e3dge_full_runner.py:286-297):
    feat_2d [...,257] (2-D-aligned features | visibility mask), feat_3d [...,256], points [...,3] world space."""
    return torch.cat([fuse_sft_mlp(feat_2d, feat_3d, sd), pos_encoding(points)], -1)


def tex_modulation(feats, sd, key=TEX):
    """feats [...,301] -> (alpha, beta) [...,256] each — volume_renderer.py:327-336."""
    m = resnet_block_fc(feats, sd, key)
    return m[..., :256], m[..., 256:]


def local_tex_modulation(feat_2d, feat_3d, points, sd):
    return tex_modulation(local_feats(feat_2d, feat_3d, points, sd), sd)
