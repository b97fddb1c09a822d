"""-m gpu: the fused epilogue + all-gather kernel (cvpp_detection_epilogue_allgather) on ONE GPU, with the
peer buffers emulated by ordinary device buffers: every destination must receive exactly the rows and counts
that cvpp_detection_epilogue produces, at the writing rank's slot."""
import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu

from computervision.pytorch_b200 import ops  # noqa: E402

DEV = "cuda:0"


@pytest.mark.parametrize("layout,width", [(ops.ROWS_FULL, 7), (ops.ROWS_YOLOV8, 6)])
def test_allgather_epilogue_matches_plain_epilogue(layout, width):
    pred = torch.from_numpy(synth.yolov8_pred(3, 4, 8400, nc=80)).to(DEV)
    det = ops.sort_nms(ops.pred_filter(pred, 80, 0.05), 0.7, max_det=50, max_nms=30000)
    table = ops.correct_boxes_params([(480, 640)] * 4, (640, 640), True, DEV)
    ref = ops.detection_epilogue(det, layout, ops.BOX_NORMALISE_CORRECT, table, packed=True)
    world, B, md = 3, 4, 50
    n_rows = world * B * md * width
    bufs = [torch.full((n_rows + world * B,), -1.0, device=DEV) for _ in range(world)]
    for rank in range(world):
        ops.detection_epilogue_allgather(det, layout, [b.data_ptr() for b in bufs], rank, ops.BOX_NORMALISE_CORRECT, table)
    torch.cuda.synchronize()
    per = B * md * width
    for b in bufs:
        for rank in range(world):
            assert torch.equal(b[rank * per:(rank + 1) * per], ref[:per])
            assert torch.equal(b[n_rows + rank * B: n_rows + (rank + 1) * B], ref[per:])
    assert np.array_equal(ref[per:].cpu().numpy(), np.minimum(det.count.cpu().numpy(), md).astype(np.float32))


def test_bad_peer_list_is_rejected():
    pred = torch.from_numpy(synth.yolov8_pred(3, 1, 2100, nc=20)).to(DEV)
    det = ops.sort_nms(ops.pred_filter(pred, 20, 0.05), 0.7, max_det=10)
    with pytest.raises(RuntimeError):
        ops.detection_epilogue_allgather(det, ops.ROWS_FULL, [0], 0)
    with pytest.raises(RuntimeError):
        ops.detection_epilogue_allgather(det, ops.ROWS_FULL, [1] * 17, 0)


@pytest.mark.parametrize("layout,width,mode", [(ops.ROWS_FULL, 7, ops.BOX_KEEP), (ops.ROWS_YOLOV8, 6, ops.BOX_NORMALISE_CORRECT),
                                               (ops.ROWS_VOC, 6, ops.BOX_NORMALISE_CORRECT)])
def test_compact_epilogue_is_the_padded_rows_without_the_padding(layout, width, mode):
    """cvpp_detection_epilogue_compact: image b's rows at row_offset[b], bit-identical to the padded epilogue's rows;
    offsets = exclusive scan of min(count, max_out); the overflow flag and the truncation at the row capacity."""
    B, md = 7, 300
    pred = torch.from_numpy(synth.yolov8_pred(11, B, 8400, nc=80)).to(DEV)
    pred[3, 4:] = 0.0                                                     # an image without detections
    det = ops.sort_nms(ops.pred_filter(pred, 80, 0.01), 0.7, max_det=md, max_nms=30000)
    table = ops.correct_boxes_params([(480, 640), (375, 500)] * 3 + [(640, 640)], (640, 640), True, DEV)
    ref = ops.detection_epilogue(det, layout, mode, table if mode != ops.BOX_KEEP else None)
    counts = np.minimum(det.count.cpu().numpy(), md)
    total = int(counts.sum())
    assert counts[3] == 0 and total > 300
    for cap in (total, total + 5, total - 100):
        rows, off, ovf = ops.detection_epilogue_compact(det, layout, cap, mode, table if mode != ops.BOX_KEEP else None)
        torch.cuda.synchronize()
        off = off.cpu().numpy()
        assert np.array_equal(off, np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
        assert int(ovf.item()) == int(total > cap)
        for b in range(B):
            lo, hi = int(off[b]), min(int(off[b + 1]), cap)
            if hi > lo:
                assert torch.equal(rows[lo:hi, :width], ref[b, :hi - lo])


class _FakeGather:
    """distributed.PeerGather's surface with the peers emulated by local buffers (one GPU, no process group)."""

    def __init__(self, world, rank, b_local, max_det, width, depth):
        self.world, self.rank, self.depth = world, rank, depth
        self.n_rows = world * b_local * max_det * width
        self.slot_elems = (self.n_rows + world * b_local + 3) & ~3
        self.bufs = [torch.full((depth * self.slot_elems,), -1.0, device=DEV) for _ in range(world)]

    def peer_ptrs(self, i):
        return [b.data_ptr() + 4 * (i % self.depth) * self.slot_elems for b in self.bufs]

    def multicast_ptr(self, i):
        return 0


@pytest.mark.parametrize("fused_rows", [True, False])
def test_pipelined_postprocess_with_the_gather_in_the_slot_graph(fused_rows):
    """PipelinedPostprocess(gather=...): decode, NMS and the gather stores of a slot replay as ONE CUDA graph - with the rows
    written by the NMS kernel's own CTAs (fused_rows, cvpp_yolov8_postprocess_gather) or by the epilogue kernel behind a
    programmatic dependent launch.  Every destination buffer must hold, at this rank's place in
    the slot, exactly the rows the plain epilogue produces from the slot's detections - for every slot, replayed twice."""
    B, A, nc, md, depth, world, rank = 2, 8400, 80, 40, 3, 3, 1
    sets = []
    for s in range(depth):
        lv = synth.yolov8_head(50 + s, B=B, nc=nc, clustered=True)
        sets.append(ops.make_levels([torch.from_numpy(np.ascontiguousarray(x)).to(DEV) for x in lv], (8.0, 16.0, 32.0)))
    fake = _FakeGather(world, rank, B, md, 7, depth)
    pipe = ops.PipelinedPostprocess(B, A, nc, torch.device(DEV), sets, 0.05, 0.7, max_det=md, gather=fake, fused_rows=fused_rows)
    per = B * md * 7
    for turn in range(2 * depth):
        slot = pipe.next_slot
        det = pipe.submit(gather=True)
        pipe.join()
        torch.cuda.synchronize()
        ref = ops.detection_epilogue(det, ops.ROWS_FULL, packed=True)
        assert int(det.count.min()) > 0
        off = slot * fake.slot_elems
        for buf in fake.bufs:
            assert torch.equal(buf[off + rank * per: off + (rank + 1) * per], ref[:per])
            assert torch.equal(buf[off + fake.n_rows + rank * B: off + fake.n_rows + (rank + 1) * B], ref[per:])
            other = (rank + 1) % world
            assert bool((buf[off + other * per: off + (other + 1) * per] == -1.0).all())   # nobody else's place is touched
    with pytest.raises(ValueError):
        ops.PipelinedPostprocess(B, A, nc, torch.device(DEV), sets, 0.05, 0.7, max_det=md).submit(gather=True)
