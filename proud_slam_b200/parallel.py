"""Ray / keyframe data parallelism over the GPUs of one box.

The reference is single-GPU.  Here every rank holds a replica of the map (octree, vertex table,
embeddings) and of the decoder, renders its own ray batch, and two exchanges make the result equal
to one big batch (up to fp32 re-association):

1. before backward, the loss couples all rays through global means and counts
   (``src/criterion.py:37-50, 96-112``): each rank's raw sums (16 doubles) are all-gathered and
   closed identically on every rank (``pslam_loss_finalize``);
2. after backward, one sum-all-reduce of a single flat fp32 buffer ``[E*16 | decoder]`` that the
   kernels scattered their gradients straight into (no staging copy before the collective).

One known difference from ONE big padded batch: the reference's first-sign-change search runs over rows padded to the batch's
longest ray with sdf = 1 (``render_helpers.py:510-545``).  A rank's (or chunk's) OWN longest ray has no pad locally, so if its
last sdf is negative and it has no earlier sign change it keeps ``z_min = z[0]`` where the big batch (longer rows) would take its
last sample; the loss closure itself uses the global row length.  At most one ray per rank and iteration; exchanging the row
length before compositing would cost a third synchronisation point per step.

``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the plumbing.  On the GPUs of one box both exchanges run
inside ``pslam_render_step`` over NVLink peer memory (``PeerExchange`` below, ``csrc/peer.cu``): the loss kernel stores its raw
sums into every peer and waits for theirs, and a two-shot kernel all-reduces the flat buffer -- no host-side collective in
the iteration.  ``DataParallelStep`` is the NCCL form of the same protocol (any transport, also the gloo tests).
"""
import ctypes as C
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block of ``n`` items owned by ``rank`` (first ranks take the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_keyframes(frame_ids: Sequence[int], rank: int, world: int) -> List[int]:
    """Keyframes of this rank when there are at least as many keyframes as GPUs
    (matches the reference's per-frame ray assembly, render_helpers.py:620-646)."""
    lo, hi = shard_bounds(len(frame_ids), rank, world)
    return list(frame_ids[lo:hi])


def shard_rays(tensors: Sequence[torch.Tensor], rank: int, world: int):
    """Contiguous ray block of the concatenated batch (dim 0 = rays)."""
    lo, hi = shard_bounds(tensors[0].shape[0], rank, world)
    return [t[lo:hi].contiguous() for t in tensors]


class FlatGrads:
    """One flat fp32 buffer ``[E*16 | decoder params]`` with views for the kernels to write into."""

    @staticmethod
    def numel(emb: torch.Tensor, dec_params: Sequence[torch.Tensor]) -> int:
        pad4 = lambda k: (k + 3) // 4 * 4
        return pad4(emb.numel()) + sum(pad4(p.numel()) for p in dec_params)

    def __init__(self, emb: torch.Tensor, dec_params: Sequence[torch.Tensor], flat: torch.Tensor = None):
        """``flat``: storage to use (e.g. ``PeerExchange.flat``, peer-mapped); allocated here when None."""
        pad4 = lambda k: (k + 3) // 4 * 4      # every segment starts 16-byte aligned (vector reductions)
        n = self.numel(emb, dec_params)
        if flat is not None and (flat.numel() != n or flat.dtype != torch.float32 or flat.device != emb.device):
            raise RuntimeError(f"flat gradient buffer must hold {n} float32 on {emb.device}")
        self.flat = torch.zeros(n, dtype=torch.float32, device=emb.device) if flat is None else flat
        self.g_emb = self.flat[: emb.numel()].view_as(emb)
        self.g_dec, off = [], pad4(emb.numel())
        for p in dec_params:
            self.g_dec.append(self.flat[off: off + p.numel()].view_as(p))
            off += pad4(p.numel())

    def zero_(self):
        self.flat.zero_()


class PeerExchange:
    """Peer-mapped buffers of one rank for the in-kernel exchanges (``pslam_peer_t``): the flat gradient buffer and the small
    exchange area, allocated as torch symmetric memory (CUDA VMM allocations every rank of the group maps) and rendezvoused
    over the process group.  ``bind(pipe)`` puts the table into a bound ``RenderPipeline``: from then on ``pipe.step()`` is a
    whole data-parallel iteration.  Collective: every rank of ``group`` must construct it (same ``flat_numel``)."""

    def __init__(self, flat_numel: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        group = dist.group.WORLD if group is None else group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"PeerExchange supports up to {_lib.MAX_PEERS} ranks (one box)")
        if flat_numel % 4:
            raise RuntimeError("flat_numel must be a multiple of 4 (FlatGrads.numel pads its segments)")
        lib = _lib.lib()
        nsync = int(lib.pslam_peer_sync_bytes())
        self.flat = symm_mem.empty(int(flat_numel), dtype=torch.float32, device=device)
        self.sync = symm_mem.empty((nsync + 7) // 8, dtype=torch.int64, device=device)
        nstage = int(lib.pslam_peer_stage_bytes(int(flat_numel), self.world))
        self.stage = symm_mem.empty((nstage + 15) // 16 * 4, dtype=torch.int32, device=device) if nstage > 0 else None
        self.flat.zero_()
        self.sync.zero_()
        if self.stage is not None:
            self.stage.zero_()
        torch.cuda.synchronize(device)
        self._h_flat = symm_mem.rendezvous(self.flat, group)
        self._h_sync = symm_mem.rendezvous(self.sync, group)
        self._h_stage = symm_mem.rendezvous(self.stage, group) if self.stage is not None else None
        dist.barrier(group)                    # every rank's exchange area is zero before anybody raises a flag in it
        self.table = _lib.PeerT()
        self.table.world, self.table.rank, self.table.flat_count = self.world, self.rank, int(flat_numel)
        for q in range(self.world):
            self.table.sync[q] = int(self._h_sync.buffer_ptrs[q])
            self.table.flat[q] = int(self._h_flat.buffer_ptrs[q])
            self.table.stage[q] = int(self._h_stage.buffer_ptrs[q]) if self._h_stage is not None else None
        self.fail = torch.zeros(1, dtype=torch.int32, device=device)
        self._lib, self._stream_ptr, self.device = lib, _lib.stream_ptr, torch.device(device)

    def bind(self, pipe, allreduce=True):
        """Installs the table in ``pipe.args`` (after ``pipe.bind``); allreduce=False: loss closure only."""
        C.memmove(C.byref(pipe.args.peer), C.byref(self.table), C.sizeof(self.table))
        if not allreduce:
            for q in range(self.world):
                pipe.args.peer.flat[q] = None
        return pipe

    def allreduce(self):
        """Sum of every rank's ``flat`` into every rank's ``flat`` (the kernel ``pslam_render_step`` ends with, on its own)."""
        from . import _lib
        _lib.check(self._lib.pslam_peer_allreduce(C.byref(self.table), _lib.ptr(self.fail), self._stream_ptr(self.device)), "pslam_peer_allreduce")


class DataParallelStep:
    """One mapping iteration across ranks.  ``pipe`` needs ``sample() / forward() / backward() /
    finalize_loss(rows)`` and a ``loss_raw`` tensor of 16 float64 (``RenderPipeline`` bound with
    ``defer_loss=True``)."""

    def __init__(self, pipe, grads: FlatGrads, group=None):
        self.pipe, self.grads, self.group = pipe, grads, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rows = torch.zeros(self.world, 16, dtype=torch.float64, device=pipe.loss_raw.device)

    def __call__(self):
        self.grads.zero_()
        self.pipe.sample()
        self.pipe.forward()
        if self.world > 1:
            dist.all_gather_into_tensor(self.rows.view(-1), self.pipe.loss_raw.view(-1), group=self.group)
        else:
            self.rows.copy_(self.pipe.loss_raw.view(1, 16))
        self.pipe.finalize_loss(self.rows)
        self.pipe.backward()
        if self.world > 1:
            dist.all_reduce(self.grads.flat, group=self.group)

    def check(self):
        """Overflow / no-hit flags of all iterations since the last call (one host sync; ``RenderPipeline.check``)."""
        return self.pipe.check() if hasattr(self.pipe, "check") else None


class ChunkedStep:
    """One mapping iteration over a ray batch larger than one launch's workspace (``pslam_wgrad_ws_bytes_w`` is 4.2 kB per
    sample of capacity, sized for either decoder build: 2^20 rays at 32 samples of capacity per ray fill 132 GiB of a
    180 GB GPU; ``bench.py --chunks`` measures 2^21 / 2^22 rays this way).  The batch is cut into chunks that go through the same two-exchange
    protocol as the ranks of ``DataParallelStep``, as *virtual ranks* of one device:

    pass 1  every chunk is rendered with the loss deferred and leaves its 16 raw sums;
            the sums of all chunks (and, with ``torch.distributed``, of all ranks) close the loss once;
    pass 2  every chunk is rendered again (same seed, hence the same samples) and runs its backward; the
            gradients of the embeddings and of the decoder accumulate in the flat buffer (the kernels add).

    ``bind_chunk(k)`` binds chunk ``k``'s rays and targets to ``pipe`` (``defer_loss=True``, a seed that depends only on
    ``k`` and the iteration); ``after_backward(k)``, if given, runs after chunk ``k``'s backward while its per-ray
    gradients (``pipe.g_rays_o / g_rays_d``) are still in place.  Costs one extra forward per chunk."""

    def __init__(self, pipe, grads: FlatGrads, n_chunks: int, bind_chunk, group=None, after_backward=None):
        self.pipe, self.grads, self.n, self.bind_chunk, self.group = pipe, grads, int(n_chunks), bind_chunk, group
        self.after_backward = after_backward
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = pipe.loss_raw.device
        self.local = torch.zeros(self.n, 16, dtype=torch.float64, device=dev)
        self.rows = torch.zeros(self.world * self.n, 16, dtype=torch.float64, device=dev)

    def __call__(self):
        self.grads.zero_()
        for k in range(self.n):
            self.bind_chunk(k)
            self.pipe.sample()
            self.pipe.forward()
            self.local[k].copy_(self.pipe.loss_raw.view(16))
        if self.world > 1:
            dist.all_gather_into_tensor(self.rows.view(-1), self.local.view(-1), group=self.group)
        else:
            self.rows.copy_(self.local)
        for k in range(self.n):
            self.bind_chunk(k)
            self.pipe.sample()
            self.pipe.forward()
            self.pipe.finalize_loss(self.rows)
            self.pipe.backward()
            if self.after_backward is not None:
                self.after_backward(k)
        if self.world > 1:
            dist.all_reduce(self.grads.flat, group=self.group)
