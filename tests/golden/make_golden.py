"""Generates tests/golden/*.npz from the REAL reference (build container only).

    python tests/golden/make_golden.py

Runs the reference's own ``render_rays`` + ``Criterion`` + ``backward``
(imported unmodified from /root/reference/src via oracle/ref_import.py, with
the CUDA-only ``grid`` kernels supplied by oracle/grid_oracle.c) on a small
synthetic scene and stores inputs, the recorded sampling noise, and every
output and gradient.  The reference has no golden vectors of its own
(SURVEY.md section 4); these are the fixtures that pin oracle/ and, through
it, the CUDA path.  /root/reference cannot travel to the GPU box, the .npz can.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import ref_import  # noqa: E402
from proud_slam_b200 import scene as sc  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CRIT = dict(rgb_weight=0.5, depth_weight=1.0, sdf_weight=5000.0, fs_weight=10.0, sdf_truncation=0.1)


def run_case(name, scene_kind, n_frames, rays_per_frame, num_embeddings, tracking, width=128):
    ref = ref_import.load()
    s = sc.make_scene(scene_kind)
    oc = oracle.Octree(s.grid_dim)
    oc.insert(s.voxels)
    v, c, f = oc.get_centres_and_children()
    ms = sc.map_states_from_flat(v, c, f, s.voxel_size, num_embeddings=max(num_embeddings, v.shape[0]))
    torch.manual_seed(0)
    dec = ref.nrgbd.Decoder(depth=2, width=width, in_dim=16, skips=[], embedder="none")
    # biases/weights a bit larger than default so colour/sdf are not degenerate
    params = list(dec.parameters())
    rays_o, rays_d, rgb, depth = sc.sample_batch(s, list(range(n_frames)), rays_per_frame, seed=3)
    # perturb the target depth a little so depth/sdf losses are non-trivial
    depth = depth * (1.0 + 0.01 * torch.randn(depth.shape, generator=torch.Generator().manual_seed(4)))
    rays_o = rays_o.clone().requires_grad_(True)
    rays_d = rays_d.clone().requires_grad_(True)
    ref.recorder.noise_chunks.clear()
    torch.manual_seed(7)
    out = ref.render_helpers.render_rays(rays_o, rays_d, ms, dec, None, 0.1 * s.voxel_size, s.voxel_size,
                                         CRIT["sdf_truncation"], 10, 10.0, return_raw=True)
    noise = torch.cat(ref.recorder.noise_chunks, 1)
    args = types.SimpleNamespace(criteria=CRIT, data_specs=dict(max_depth=10.0))
    crit = ref.criterion.Criterion(args)
    if tracking:
        hit = out["ray_mask"].view(-1)
        out["ray_mask"] = hit
        loss, parts = crit(out, (rgb[0], depth[0]), weight_depth_loss=True)
    else:
        loss, parts = crit(out, (rgb, depth))
    loss.backward()
    # intermediate tensors: rerun the reference pre-processing with the same noise is not
    # possible (it draws its own), so store what the recorder and outputs give us.
    inter, hits = ref.voxel_helpers.ray_intersect_vox(rays_o.detach(), rays_d.detach(), ms["voxel_center_xyz"],
                                                      ms["voxel_structure"], s.voxel_size, 10, 10.0)
    data = dict(
        voxel_size=np.float32(s.voxel_size), grid_dim=np.int32(s.grid_dim), voxels=s.voxels,
        centres=ms["voxel_center_xyz"].numpy(), structure=ms["voxel_structure"].numpy(),
        vertex_idx=ms["voxel_vertex_idx"].numpy(), emb=ms["voxel_vertex_emb"].detach().numpy(),
        rays_o=rays_o.detach().numpy(), rays_d=rays_d.detach().numpy(), rgb=rgb.numpy(), depth=depth.numpy(),
        noise=noise.numpy(), tracking=np.int32(tracking), step_size=np.float32(0.1 * s.voxel_size),
        truncation=np.float32(CRIT["sdf_truncation"]), max_distance=np.float32(10.0), max_depth=np.float32(10.0),
        crit_weights=np.array([CRIT["rgb_weight"], CRIT["depth_weight"], CRIT["fs_weight"], CRIT["sdf_weight"]], np.float32),
        hit_idx=inter["intersected_voxel_idx"].numpy(), hit_min=inter["min_depth"].numpy(),
        hit_max=inter["max_depth"].numpy(), hits=hits.numpy(),
        out_weights=out["weights"].detach().numpy(), out_color=out["color"].detach().numpy(),
        out_depth=out["depth"].detach().numpy(), out_z_vals=out["z_vals"].numpy(),
        out_sdf=out["sdf"].detach().numpy(), out_ray_mask=out["ray_mask"].numpy(), out_raw=out["raw"].numpy(),
        loss=np.float32(loss.item()),
        loss_parts=np.array([parts["color_loss"], parts["depth_loss"], parts["fs_loss"], parts["sdf_loss"]], np.float32),
        g_emb=ms["voxel_vertex_emb"].grad.numpy(), g_rays_o=rays_o.grad.numpy(), g_rays_d=rays_d.grad.numpy(),
    )
    for i, p in enumerate(params):
        data[f"dec_{i}"] = p.detach().numpy()
        data[f"g_dec_{i}"] = p.grad.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **data)
    print(name, "R", rays_o.shape[1], "hit", int(hits.sum()), "S", out["z_vals"].shape, "loss", loss.item(),
          "%.0f kB" % (os.path.getsize(path) / 1e3))


def se3_kat():
    """Known-answer check the reference carries itself (se3pose.py:103-113)."""
    ref = ref_import.load()
    before = torch.tensor([[-0.955421, 0.119616, -0.269932, 2.655830],
                           [0.295248, 0.388339, -0.872939, 2.981598],
                           [0.000408, -0.913720, -0.406343, 1.368648],
                           [0.0, 0.0, 0.0, 1.0]])
    pose = ref.se3pose.OptimizablePose.from_matrix(before)
    np.savez(os.path.join(HERE, "se3_kat.npz"), before=before.numpy(), data=pose.data.detach().numpy(),
             rotation=pose.rotation().detach().numpy(), after=pose.matrix().detach().numpy())


if __name__ == "__main__":
    assert ref_import.available(), "needs /root/reference"
    run_case("mapping_tiny", "tiny", 2, 192, 512, tracking=False)
    run_case("tracking_tiny", "tiny", 1, 256, 512, tracking=True)
    run_case("mapping_tiny_w256", "tiny", 2, 96, 512, tracking=False, width=256)
    # BASELINE.json configs[0]: 2048-ray mapping step on the 0.2 m Replica-shaped octree, default-init decoder
    run_case("mapping_replica_small", "replica_small", 2, 1024, 0, tracking=False)
    se3_kat()
