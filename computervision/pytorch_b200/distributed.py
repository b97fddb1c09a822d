"""Image sharding and the detection all-gather.

Every image is independent, so a batch is split into contiguous image ranges, one per rank (one process per
GPU), and no collective touches the data path.  The only exchange is the all-gather of the padded detection
rows `(B_local, max_det, W)` + per-image counts that an evaluation needs.  The helpers are backend-agnostic
(NCCL on GPUs; gloo on CPU tensors for the host-logic tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `rank`: the first n % world ranks get one extra item."""
    if world_size < 1 or not (0 <= rank < world_size) or n_items < 0:
        raise ValueError(f"bad shard request: n={n_items} world={world_size} rank={rank}")
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world_size: int) -> List[int]:
    return [shard_range(n_items, world_size, r)[1] - shard_range(n_items, world_size, r)[0] for r in range(world_size)]


def shard_batch(tensors: Sequence[torch.Tensor], world_size: int, rank: int) -> List[torch.Tensor]:
    """The rank's slice (a view, no copy) of every per-image tensor of a global batch."""
    n = int(tensors[0].shape[0])
    s, e = shard_range(n, world_size, rank)
    return [t[s:e] for t in tensors]


def gather_detections(rows: torch.Tensor, counts: torch.Tensor, n_global: Optional[int] = None,
                      group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of per-image detection rows.

    rows (B_local, max_det, W) float32 and counts (B_local,) int32 of this rank's shard (shard_range of
    n_global images; n_global defaults to world * B_local) -> (rows_all (n_global, max_det, W),
    counts_all (n_global,)) in global image order on every rank.  Uneven shards are padded to the largest
    shard for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return rows, counts
    world = dist.get_world_size(group)
    b_local = int(rows.shape[0])
    if n_global is None:
        n_global = world * b_local
    sizes = shard_sizes(n_global, world)
    if sizes[dist.get_rank(group)] != b_local:
        raise ValueError(f"rank holds {b_local} images, shard_range says {sizes[dist.get_rank(group)]}")
    b_max = max(sizes)
    if b_local < b_max:  # pad to a common shape
        pad_r = rows.new_zeros((b_max - b_local,) + tuple(rows.shape[1:]))
        rows = torch.cat((rows, pad_r), 0)
        counts = torch.cat((counts, counts.new_zeros((b_max - b_local,))), 0)
    rows, counts = rows.contiguous(), counts.contiguous()
    rows_all = rows.new_empty((world * b_max,) + tuple(rows.shape[1:]))
    counts_all = counts.new_empty((world * b_max,))
    if rows.is_cuda:
        dist.all_gather_into_tensor(rows_all, rows, group=group)
        dist.all_gather_into_tensor(counts_all, counts, group=group)
    else:  # gloo: list form
        dist.all_gather(list(rows_all.chunk(world, 0)), rows, group=group)
        dist.all_gather(list(counts_all.chunk(world, 0)), counts, group=group)
    if all(s == b_max for s in sizes):
        return rows_all, counts_all
    keep = torch.cat([torch.arange(r * b_max, r * b_max + s) for r, s in enumerate(sizes)]).to(rows_all.device)
    return rows_all.index_select(0, keep), counts_all.index_select(0, keep)


def gather_packed(packed: torch.Tensor, group=None) -> torch.Tensor:
    """ONE all-gather of the flat [rows | counts] buffer of ops.detection_epilogue(packed=True) (equal shards):
    returns (world, len(packed))."""
    if not dist.is_available() or not dist.is_initialized():
        return packed.reshape(1, -1)
    world = dist.get_world_size(group)
    out = packed.new_empty((world, packed.numel()))
    if packed.is_cuda:
        dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
    else:
        dist.all_gather(list(out.unbind(0)), packed.contiguous(), group=group)
    return out


def unpack_detections(gathered: torch.Tensor, b_local: int, max_det: int, width: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(world, B*max_det*W + B) -> (rows (world*B, max_det, W), counts (world*B,) int32) in global image order."""
    world = gathered.shape[0]
    n_rows = b_local * max_det * width
    rows = gathered[:, :n_rows].reshape(world * b_local, max_det, width)
    counts = gathered[:, n_rows:].reshape(world * b_local).to(torch.int32)
    return rows, counts


class PeerGather:
    """Gather buffers for cvpp_detection_epilogue_allgather: one symmetric-memory allocation per rank, mapped
    into every peer over NVLink (torch.distributed._symmetric_memory), `depth` alternating slots so that the
    stores of step k+1 never land in the buffer a consumer of step k may still be reading.

        pg = PeerGather(b_local, max_det, width, device)          # collective
        ops.detection_epilogue_allgather(det, ops.ROWS_FULL, pg.peer_ptrs(i), pg.rank)
        pg.barrier(i)                                             # stream-ordered cross-rank barrier
        rows, counts = pg.view(i)                                 # (world*B, max_det, W), (world*B,) int32
    """

    def __init__(self, b_local: int, max_det: int, width: int, device, depth: int = 2, group=None):
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.b_local, self.max_det, self.width, self.depth = int(b_local), int(max_det), int(width), int(depth)
        self.n_rows = self.world * self.b_local * self.max_det * self.width
        self.slot_elems = (self.n_rows + self.world * self.b_local + 3) & ~3       # 16-byte aligned slots
        self.buf = symm.empty((self.depth * self.slot_elems,), dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        self._ptrs = [int(p) for p in self.handle.buffer_ptrs]
        # NVSwitch multicast mapping of the same buffer (0 when the box / driver has no NVLS support)
        try:
            self._mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        except Exception:
            self._mc = 0

    def peer_ptrs(self, i: int) -> List[int]:
        off = 4 * (i % self.depth) * self.slot_elems
        return [p + off for p in self._ptrs]

    def multicast_ptr(self, i: int) -> int:
        """Multicast address of slot i (0 = not available: use peer_ptrs)."""
        return self._mc + 4 * (i % self.depth) * self.slot_elems if self._mc else 0

    def barrier(self, i: int = 0) -> None:
        self.handle.barrier(channel=i % self.depth)

    def view(self, i: int) -> Tuple[torch.Tensor, torch.Tensor]:
        s = self.buf[(i % self.depth) * self.slot_elems:][: self.n_rows + self.world * self.b_local]
        rows = s[: self.n_rows].reshape(self.world * self.b_local, self.max_det, self.width)
        return rows, s[self.n_rows:].to(torch.int32)

    def view_counts_f32(self, i: int) -> torch.Tensor:
        """The (world*B,) per-image counts of slot i as stored (fp32), a view - no conversion kernel."""
        s = self.buf[(i % self.depth) * self.slot_elems:]
        return s[self.n_rows: self.n_rows + self.world * self.b_local]


def split_rows(rows_all: torch.Tensor, counts_all: torch.Tensor) -> List[torch.Tensor]:
    """Per-image (n_i, W) views of the gathered rows (one host read of the counts)."""
    cap = rows_all.shape[1]
    return [rows_all[i, :min(int(n), cap)] for i, n in enumerate(counts_all.tolist())]
