"""Live pin of the oracle and of the state_dict contract against the REAL reference.
Only runs where /root/reference exists (the build container); runs the reference in a
subprocess because its top-level package is called `project`, like our shim."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle import ref_harness as rh

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not present")

_SCRIPT = r"""
import json, sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests')
from oracle import ref_harness as rh, params as P, stylesdf_oracle as O
from helpers import generator_state_dict_spec, rel_linf
ref = rh.load_reference()
res = {}
for size, r, local in ((256, 64, False), (1024, 64, False)):
    G = ref.stylesdf_model.G_pred_latents(rh.model_opt(size=size, renderer_spatial_output_dim=r),
                                          rh.rendering_opt())
    got = {k: tuple(v.shape) for k, v in G.state_dict().items()}
    want = {k: tuple(v) for k, v in generator_state_dict_spec(size, r, local).items()}
    res['keys_%%d' %% size] = (got == want)
# live forward: reference vs oracle, random-init weights of the reference itself
torch.manual_seed(3)
G = ref.stylesdf_model.G_pred_latents(rh.model_opt(size=64, renderer_spatial_output_dim=16),
                                      rh.rendering_opt()).eval()
sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
inp = P.make_inputs(5, 2, G.decoder.n_latent, 16)
with torch.no_grad():
    a = G([inp['w'], inp['w_dec']], inp['cam_poses'], inp['focal'], inp['near'], inp['far'],
          input_is_latent=True, randomize_noise=False, return_xyz=True, return_sdf=True)
    b = O.generator_forward(sd, inp['w'], inp['w_dec'], inp['cam_poses'], inp['focal'],
                            inp['near'], inp['far'], res=16, n_samples=24)
res['err'] = {k: rel_linf(b[k], a[k]) for k in ('features', 'gen_thumb_imgs', 'sdf', 'hit_prob',
                                                 'xyz', 'depth', 'dists', 'points', 'gen_imgs')}
print('RESULT ' + json.dumps(res))
"""


def test_oracle_and_contract_against_live_reference():
    r = subprocess.run([sys.executable, "-c", _SCRIPT % {"root": ROOT}], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    res = json.loads(line[7:])
    assert res["keys_256"] and res["keys_1024"]
    for k, e in res["err"].items():
        assert e < 2e-5, (k, e)


def test_local_feature_query_oracle_matches_the_live_reference():
    """oracle/local_query_oracle.py against the reference's HGPIFuNetGAN.query executed here (fresh random
    cases, not the committed fixture): SURVEY.md 8f row 1."""
    script = r"""
import json, sys, torch
sys.path.insert(0, %(root)r)
from oracle import gen_golden_local_query as gen, local_query_oracle as LQ
res = {}
for i, args in enumerate(((101, 2, 8, 10, 14, 300, True), (102, 1, 12, 16, 9, 200, False))):
    feat, pts, calibs = gen.make_case(*args)
    ref = gen.run(feat, pts, calibs)
    got = LQ.local_feature_query(pts, calibs, feat)
    res[str(i)] = {k: float((got[k] - ref[k]).abs().max() / max(float(ref[k].abs().max()), 1.0))
                   for k in ("proj_xy", "depth", "feats")}
    res[str(i)]["in_img"] = bool((got["in_img"] == ref["in_img"]).all())
print("RESULT" + json.dumps(res))
""" % {"root": ROOT}
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.split("RESULT")[-1])
    for case in res.values():
        assert case["in_img"]
        for k in ("proj_xy", "depth", "feats"):
            assert case[k] < 2e-6, (k, case[k])


def test_second_order_eikonal_gradient_of_the_oracle_matches_the_live_reference():
    """The reference's own `get_eikonal_term` (autograd.grad(..., create_graph=True), volume_renderer.py:796-802)
    inside its renderer, an eikonal loss on it (losses/gan_loss.py) and its gradient with respect to the w+
    latents — against oracle.eikonal_term differentiated the same way."""
    script = r"""
import json, sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + '/tests')
from oracle import ref_harness as rh, params as P, stylesdf_oracle as O
from helpers import rel_linf
ref = rh.load_reference()
torch.manual_seed(4)
G = ref.stylesdf_model.G_pred_latents(rh.model_opt(size=64, renderer_spatial_output_dim=8),
                                      rh.rendering_opt(N_samples=6), full_pipeline=False).eval()
sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
inp = P.make_inputs(6, 2, 2, 8)
w = inp['w'].clone().requires_grad_(True)
out = G.renderer(inp['cam_poses'], inp['focal'], inp['near'], inp['far'], styles=w, return_eikonal=True)
eik = out['eikonal_term']
loss = ((eik.norm(dim=-1) - 1) ** 2).mean()
g_ref, = torch.autograd.grad(loss, [w])
w2 = inp['w'].clone().requires_grad_(True)
e2 = O.eikonal_term(sd, out['points'].detach().reshape(2, -1, 3), w2).reshape(eik.shape)
g_or, = torch.autograd.grad(((e2.norm(dim=-1) - 1) ** 2).mean(), [w2])
print('RESULT' + json.dumps({'eik': rel_linf(e2.detach(), eik.detach()), 'grad': rel_linf(g_or, g_ref),
                             'gmax': float(g_ref.abs().max())}))
""" % {"root": ROOT}
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    res = json.loads(out.stdout.split("RESULT")[-1])
    assert res["gmax"] > 0 and res["eik"] < 2e-5 and res["grad"] < 1e-4, res
