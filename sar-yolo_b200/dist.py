"""Multi-GPU plumbing: batch / tile sharding and the one real exchange step.

Images are independent, so ranks post-process disjoint slices of the batch with no data-path
collective.  The only exchange is an all-gather of the per-image detection counts and the padded
`(B/W, max_det, row_len)` detection rows — the fused gather kernel (K5) writes its rows directly into
this rank's slot of the all-gather buffer, so there is no staging copy before NCCL.  For sliced
inference the tiles of a frame are kept on one rank whenever possible so the cross-tile merge is
rank-local; otherwise tile detections are all-gathered and every rank merges the frames it owns.
Works with backend "nccl" (GPU) and "gloo" (CPU tensors, used by the host-logic tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced slice [lo, hi) of `n_items` owned by `rank` (first `n % world` ranks get one more)."""
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_frames(n_frames: int, tiles_per_frame: int, rank: int, world: int) -> Tuple[int, int]:
    """Tile range [lo, hi) for `rank` when whole frames are assigned to ranks (merge stays rank-local)."""
    f_lo, f_hi = shard_range(n_frames, rank, world)
    return f_lo * tiles_per_frame, f_hi * tiles_per_frame


def sahi_grid(frame_w: int, frame_h: int, tile: int = 640, overlap: float = 0.2) -> torch.Tensor:
    """Tile origins (T, 2) as (x0, y0) of a SAHI-style slicing grid: step = tile*(1-overlap), last
    row/column snapped back to the border (4000x3000 -> 8 x 6 = 48 tiles)."""
    step = int(tile * (1.0 - overlap))

    def axis(n):
        if n <= tile:
            return [0]
        xs = list(range(0, n - tile, step))
        xs.append(n - tile)
        return xs

    xs, ys = axis(frame_w), axis(frame_h)
    return torch.tensor([(x, y) for y in ys for x in xs], dtype=torch.float32)


def allgather_detections(out: torch.Tensor, counts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather padded detections.  `out (b, max_det, row_len)`, `counts (b,)` with the same `b` on
    every rank -> `(W*b, max_det, row_len)`, `(W*b,)` ordered by rank."""
    world = dist.get_world_size(group)
    if world == 1:
        return out, counts
    g_out = torch.empty((world * out.shape[0],) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
    g_cnt = torch.empty((world * counts.shape[0],), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(g_out, out.contiguous(), group=group)
    dist.all_gather_into_tensor(g_cnt, counts.contiguous(), group=group)
    return g_out, g_cnt


class GatherBuffer:
    """Pre-allocated all-gather destination whose slot `rank` is handed to the gather kernel as its
    output buffer: K5 writes straight into the send region and NCCL runs in place."""

    def __init__(self, per_rank_batch: int, max_det: int, row_len: int, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.per = int(per_rank_batch)
        self.out = torch.empty((self.world * self.per, max_det, row_len), dtype=torch.float32, device=device)
        self.counts = torch.zeros((self.world * self.per,), dtype=torch.int32, device=device)

    @property
    def local_out(self) -> torch.Tensor:
        return self.out[self.rank * self.per:(self.rank + 1) * self.per]

    @property
    def local_counts(self) -> torch.Tensor:
        return self.counts[self.rank * self.per:(self.rank + 1) * self.per]

    def exchange(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """In-place all-gather (send buffer = this rank's slice of the receive buffer, the layout NCCL's
        in-place all-gather expects)."""
        if self.world > 1:
            dist.all_gather_into_tensor(self.out, self.local_out, group=self.group)
            dist.all_gather_into_tensor(self.counts, self.local_counts, group=self.group)
        return self.out, self.counts


class PeerGatherBuffer:
    """All-gather destination in NVLink peer memory (torch symmetric memory): every rank holds the full
    `(W*per, max_det, row_len)` rows + `(W*per,)` counts, and every rank's gather kernel (K5) stores its rows
    straight into ALL ranks' buffers (P2P stores over NVLink/NVSwitch) — the exchange is fused into the kernel that
    produces the data; no NCCL collective, no staging copy.  `barrier()` (a device-side signal-pad barrier on the
    current stream) orders those stores before any rank reads.  Two alternating buffers let step i+1 write while a
    slow peer still reads step i (its barrier i+1 comes after its reads of step i in stream order).
    """

    def __init__(self, per_rank_batch: int, max_det: int, row_len: int, device, group=None, n_buffers: int = 2,
                 total_slots: int = None, slot_offset: int = None):
        """Default layout: every rank contributes `per_rank_batch` image slots, rank r at slot r*per.  Uneven shards
        (whole frames per rank, `shard_frames`) pass `total_slots` and this rank's `slot_offset` explicitly."""
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.per, self.max_det, self.row_len = int(per_rank_batch), int(max_det), int(row_len)
        total = self.world * self.per if total_slots is None else int(total_slots)
        self._slot_offset = self.rank * self.per if slot_offset is None else int(slot_offset)
        if self._slot_offset + self.per > total:
            raise ValueError("sarpost: this rank's slots exceed the gather buffer")
        self._bufs = []
        for _ in range(n_buffers):
            rows = symm_mem.empty((total, self.max_det, self.row_len), dtype=torch.float32, device=device)
            cnts = symm_mem.empty((total,), dtype=torch.int32, device=device)
            h_rows = symm_mem.rendezvous(rows, group)
            h_cnts = symm_mem.rendezvous(cnts, group)
            self._bufs.append((rows, cnts, h_rows, h_cnts))
        self._i = 0

    def next(self) -> "PeerGatherBuffer":
        """Advance to the other buffer (call once per step, before the fused call)."""
        self._i = (self._i + 1) % len(self._bufs)
        return self

    @property
    def rows(self) -> torch.Tensor:
        return self._bufs[self._i][0]

    @property
    def counts(self) -> torch.Tensor:
        return self._bufs[self._i][1]

    @property
    def slot_offset(self) -> int:
        return self._slot_offset

    def peer_ptrs(self):
        _, _, h_rows, h_cnts = self._bufs[self._i]
        return list(h_rows.buffer_ptrs), list(h_cnts.buffer_ptrs)

    def barrier(self) -> None:
        """All ranks' peer stores issued before this point (on their current streams) are visible after it."""
        self._bufs[self._i][2].barrier(channel=0)
