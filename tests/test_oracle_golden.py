"""Pins oracle/ against the fixtures generated from the REAL reference Python
(tests/golden/make_golden.py ran the reference's render_rays + Criterion + backward, imported
unmodified from /root/reference/src).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import render_oracle as ro
from tests import util
from tests.util import rel_err


@pytest.mark.parametrize("name", ["mapping_tiny", "tracking_tiny", "mapping_tiny_w256", "mapping_replica_small"])
def test_oracle_reproduces_reference_fixture(name):
    g = util.load_golden(name)
    ms = util.golden_map_states(g)
    dec = util.golden_decoder(g)
    t = lambda k: torch.from_numpy(g[k])
    rays_o, rays_d = t("rays_o").requires_grad_(True), t("rays_d").requires_grad_(True)
    out, loss, parts = util.oracle_step(rays_o, rays_d, t("rgb"), t("depth"), ms, dec, voxel_size=float(g["voxel_size"]),
                                        noise=t("noise"), tracking=bool(g["tracking"]))
    inter = out["_dbg"]["intersections"]
    hit = torch.from_numpy(g["hits"]).view(-1)
    assert np.array_equal(out["ray_mask"].view(-1).numpy(), g["hits"].reshape(-1))
    assert np.array_equal(inter["intersected_voxel_idx"].numpy(), g["hit_idx"][0][hit.numpy()])
    assert np.array_equal(inter["min_depth"].numpy(), g["hit_min"][0][hit.numpy()])
    assert np.array_equal(out["z_vals"].numpy(), g["out_z_vals"])            # sample depths: bit-exact
    assert np.array_equal(out["sdf"].detach().numpy() == 1.0, g["out_sdf"] == 1.0)      # pads (sdf = 1) at the same places
    for k, gk in (("sdf", "out_sdf"), ("color", "out_color"), ("depth", "out_depth"), ("weights", "out_weights")):
        assert rel_err(out[k].detach(), g[gk]) < 1e-6, k
    assert abs(float(loss) - float(g["loss"])) < 1e-6 * abs(float(g["loss"]))
    for k, ref in zip(("color_loss", "depth_loss", "fs_loss", "sdf_loss"), g["loss_parts"]):
        assert abs(float(parts[k]) - float(ref)) <= 1e-6 * max(abs(float(ref)), 1e-12)
    assert rel_err(ms["voxel_vertex_emb"].grad, g["g_emb"]) < 1e-5
    assert rel_err(rays_o.grad, g["g_rays_o"]) < 1e-5
    assert rel_err(rays_d.grad, g["g_rays_d"]) < 1e-5
    for i in range(10):
        assert rel_err(dec[i].grad, g[f"g_dec_{i}"]) < 1e-5


def test_se3_known_answer():
    """The one self-check the reference carries (src/se3pose.py:103-113): matrix -> (t, w) -> matrix."""
    g = util.load_golden("se3_kat")
    w = torch.from_numpy(g["data"][3:])
    R = ro.se3_rotation(w)
    assert np.allclose(R.numpy(), g["rotation"], atol=1e-6)
    assert np.allclose(g["after"][:3, :3], g["before"][:3, :3], atol=1e-4)
