"""Worker of tests/test_gpu_peer.py (one process per GPU, launched by torchrun): the in-kernel exchanges of csrc/peer.cu against
the NCCL form of the same protocol and against one single-GPU step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    from oracle import render_oracle as ro
    from proud_slam_b200 import scene as sc
    from proud_slam_b200.parallel import DataParallelStep, FlatGrads, PeerExchange
    from proud_slam_b200.pipeline import RenderPipeline
    from tests import util
    s, ms_cpu = util.build_scene("replica_small")
    ms = {k: v.detach().to(device).contiguous() for k, v in ms_cpu.items()}
    dec = [p.detach().to(device).contiguous() for p in ro.decoder_params(width=128, seed=1)]
    out = {}

    def bind(pipe, batch, fg, seed, defer):
        pipe.bind(batch[0], batch[1], ms, dec, voxel_size=s.voxel_size, step_size=0.1 * s.voxel_size, truncation=0.1, max_distance=10.0,
                  target_rgb=batch[2], target_depth=batch[3], noise=None, seed=seed, g_emb=fg.g_emb, g_dec=fg.g_dec, grad_rays=True,
                  defer_loss=defer)

    def batch_of(r, n=512):
        b = sc.sample_batch(s, [2 * r, 2 * r + 1], n, seed=50 + r)
        return [t[0].to(device).contiguous() for t in b]

    peer = PeerExchange(FlatGrads.numel(ms["voxel_vertex_emb"], dec), device)
    fg_peer = FlatGrads(ms["voxel_vertex_emb"], dec, flat=peer.flat)
    fg_nccl = FlatGrads(ms["voxel_vertex_emb"], dec)
    mine = batch_of(rank)
    # 1. different batches per rank: in-kernel exchanges == NCCL form, repeated (epochs advance, parities alternate)
    pipe_a = RenderPipeline(mine[0].shape[0], device, samples_per_ray=96)
    pipe_b = RenderPipeline(mine[0].shape[0], device, samples_per_ray=96)
    worst_g, worst_l = 0.0, 0.0
    for it in range(5):
        bind(pipe_a, mine, fg_peer, 10 + it, False)
        peer.bind(pipe_a)
        fg_peer.zero_()
        pipe_a.step()
        bind(pipe_b, mine, fg_nccl, 10 + it, True)
        DataParallelStep(pipe_b, fg_nccl)()
        torch.cuda.synchronize()
        pipe_a.check()
        la, lb = float(pipe_a.loss[0]), float(pipe_b.loss[0])
        worst_l = max(worst_l, abs(la - lb) / abs(lb))
        worst_g = max(worst_g, float((fg_peer.flat - fg_nccl.flat).abs().max() / fg_nccl.flat.abs().max()))
    out["loss_vs_nccl"], out["grad_vs_nccl"] = worst_l, worst_g
    # 2. all ranks hold bit-identical sums
    lo, hi = fg_peer.flat.clone(), fg_peer.flat.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["bit_identical"] = bool(torch.equal(lo, hi))
    # 3. same batch everywhere == one single-GPU step on it
    b0 = batch_of(0)
    bind(pipe_a, b0, fg_peer, 99, False)
    peer.bind(pipe_a)
    fg_peer.zero_()
    pipe_a.step()
    torch.cuda.synchronize()
    multi, lm = fg_peer.flat.clone(), float(pipe_a.loss[0])
    single = FlatGrads(ms["voxel_vertex_emb"], dec)
    pipe_c = RenderPipeline(b0[0].shape[0], device, samples_per_ray=96)
    bind(pipe_c, b0, single, 99, False)
    pipe_c.step()
    torch.cuda.synchronize()
    out["loss_vs_single"] = abs(lm - float(pipe_c.loss[0])) / abs(float(pipe_c.loss[0]))
    out["grad_vs_single"] = float((multi - single.flat).abs().max() / single.flat.abs().max())
    # 4. the all-reduce on its own (odd sizes of the slices: flat_count / 4 not divisible by the world)
    peer.flat.copy_(torch.arange(peer.flat.numel(), device=device, dtype=torch.float32) * 1e-3 + rank)
    dist.barrier()
    torch.cuda.synchronize()
    peer.allreduce()
    torch.cuda.synchronize()
    want = torch.arange(peer.flat.numel(), device=device, dtype=torch.float32) * 1e-3 * world + sum(range(world))
    out["allreduce_err"] = float((peer.flat - want).abs().max())
    out["fail_flag"] = int(peer.fail.item())
    dist.barrier()
    if rank == 0:
        print("PEER_RESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
