"""Generates tests/golden/loop_*.npz from the REAL reference loops (build container only).

    python tests/golden/make_loop_golden.py

Runs the reference's own ``bundle_adjust_frames`` and ``track_frame`` (src/variations/render_helpers.py:559-761, imported
unmodified through oracle/ref_import.py; the CUDA-only ``grid`` kernels supplied by oracle/grid_oracle.c) for a few iterations
on the tiny synthetic scene with deterministic frame stubs, and stores what is needed to replay them: the frames, every
iteration's pixel selection and sampling noise, and the state the loops leave behind (embeddings, decoder, poses, losses).
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


class StubFrame:
    """What the loops touch of src/frame.py's RGBDFrame: rays_d / rgb / depth per pixel, a pose with its own Adam, and
    sample_rays(n) -- here a recorded, seeded draw of n distinct pixels (sorted: mask-indexing order)."""

    def __init__(self, ref, scene, frame, stamp, perturb, seed):
        self.stamp = stamp
        self.rays_d = scene.rays_cam.reshape(-1, 3).clone()
        self.rgb = frame.rgb.reshape(-1, 3).clone()
        self.depth = frame.depth.reshape(-1).clone()
        pose = frame.pose.clone()
        pose[:3, 3] += torch.tensor(perturb)
        self.pose = ref.se3pose.OptimizablePose.from_matrix(pose)
        self.optim = torch.optim.Adam(self.pose.parameters(), lr=5e-3)
        self.gen = torch.Generator().manual_seed(seed)
        self.sample_mask = None
        self.masks = []

    def get_pose(self):
        return self.pose.matrix()

    def sample_rays(self, n):
        self.sample_mask = torch.randperm(self.rays_d.shape[0], generator=self.gen)[:n].sort().values
        self.masks.append(self.sample_mask.clone())


class RecordingCriterion:
    def __init__(self, inner):
        self.inner, self.losses = inner, []
        for k in ("rgb_weight", "depth_weight", "fs_weight", "sdf_weight", "truncation", "max_dpeth"):
            setattr(self, k, getattr(inner, k))

    def __call__(self, *a, **k):
        loss, parts = self.inner(*a, **k)
        self.losses.append(float(loss))
        return loss, parts


def build(ref):
    s = sc.make_scene("tiny")
    oc = oracle.Octree(s.grid_dim)
    oc.insert(s.voxels)
    v, c, f = oc.get_centres_and_children()
    ms = sc.map_states_from_flat(v, c, f, s.voxel_size, num_embeddings=max(600, v.shape[0]))
    torch.manual_seed(0)
    dec = ref.nrgbd.Decoder(depth=2, width=128, in_dim=16, skips=[], embedder="none")
    crit = RecordingCriterion(ref.criterion.Criterion(types.SimpleNamespace(criteria=CRIT, data_specs=dict(max_depth=10.0))))
    return s, ms, dec, crit


def common(s, ms, dec):
    return dict(voxel_size=np.float32(s.voxel_size), centres=ms["voxel_center_xyz"].numpy(), structure=ms["voxel_structure"].numpy(),
                vertex_idx=ms["voxel_vertex_idx"].numpy(), emb0=ms["voxel_vertex_emb"].detach().numpy().copy(),
                rays_cam=s.rays_cam.reshape(-1, 3).numpy(), **{f"dec0_{i}": p.detach().numpy().copy() for i, p in enumerate(dec.parameters())})


def make_ba(iters=4, n_rays=128):
    ref = ref_import.load()
    s, ms, dec, crit = build(ref)
    data = common(s, ms, dec)
    frames = [StubFrame(ref, s, s.frames[0], 0, [0.0, 0.0, 0.0], 11), StubFrame(ref, s, s.frames[1], 1, [0.02, -0.01, 0.015], 12)]
    for k, fr in enumerate(frames):
        data[f"rgb_{k}"], data[f"depth_{k}"], data[f"pose0_{k}"] = fr.rgb.numpy(), fr.depth.numpy(), fr.pose.data.detach().numpy().copy()
    emb = ms["voxel_vertex_emb"]
    embed_optim = torch.optim.Adam([emb], lr=1e-2)
    model_optim = torch.optim.Adam(dec.parameters(), lr=1e-2)
    ref.recorder.noise_chunks.clear()
    torch.manual_seed(7)
    ref.render_helpers.bundle_adjust_frames(frames, ms, dec, None, crit, s.voxel_size, 0.1 * s.voxel_size, N_rays=n_rays, num_iterations=iters,
                                            truncation=CRIT["sdf_truncation"], max_voxel_hit=10, max_distance=10.0,
                                            embed_optim=embed_optim, model_optim=model_optim, update_pose=True)
    assert len(ref.recorder.noise_chunks) == iters, len(ref.recorder.noise_chunks)      # one 800-ray chunk per render_rays call
    for i, nz in enumerate(ref.recorder.noise_chunks):
        data[f"noise_{i}"] = nz.reshape(-1, nz.shape[-1]).numpy()
    for k, fr in enumerate(frames):
        data[f"masks_{k}"] = torch.stack(fr.masks).numpy()
        data[f"pose1_{k}"] = fr.pose.data.detach().numpy()
    data.update(emb1=emb.detach().numpy(), losses=np.array(crit.losses, np.float32), iters=np.int32(iters), n_rays=np.int32(n_rays),
                **{f"dec1_{i}": p.detach().numpy() for i, p in enumerate(dec.parameters())})
    np.savez_compressed(os.path.join(HERE, "loop_ba_tiny.npz"), **data)
    print("loop_ba_tiny: losses", crit.losses, "pose moved", float((frames[1].pose.data.detach() - torch.from_numpy(data["pose0_1"])).abs().max()))


def make_track(iters=6, n_rays=256):
    ref = ref_import.load()
    s, ms, dec, crit = build(ref)
    data = common(s, ms, dec)
    fr = StubFrame(ref, s, s.frames[1], 1, [0.03, -0.02, 0.01], 21)
    data["rgb_0"], data["depth_0"], data["pose0_0"] = fr.rgb.numpy(), fr.depth.numpy(), fr.pose.data.detach().numpy().copy()
    ref.recorder.noise_chunks.clear()
    torch.manual_seed(9)
    pose, _, hit = ref.render_helpers.track_frame(fr.pose, fr, ms, dec, None, crit, s.voxel_size, N_rays=n_rays, step_size=0.1 * s.voxel_size,
                                                  num_iterations=iters, truncation=CRIT["sdf_truncation"], learning_rate=1e-2, max_voxel_hit=10,
                                                  max_distance=10.0, depth_variance=True)
    assert len(ref.recorder.noise_chunks) == iters
    for i, nz in enumerate(ref.recorder.noise_chunks):
        data[f"noise_{i}"] = nz.reshape(-1, nz.shape[-1]).numpy()
    data.update(masks_0=torch.stack(fr.masks).numpy(), pose1_0=pose.data.detach().numpy(), losses=np.array(crit.losses, np.float32),
                hit_mask=hit.numpy(), iters=np.int32(iters), n_rays=np.int32(n_rays))
    np.savez_compressed(os.path.join(HERE, "loop_track_tiny.npz"), **data)
    print("loop_track_tiny: losses", crit.losses, "pose moved", float((pose.data.detach() - torch.from_numpy(data["pose0_0"])).abs().max()))


if __name__ == "__main__":
    make_ba()
    make_track()
