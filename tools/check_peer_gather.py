"""2+ GPU check of the fused epilogue + NVLink peer-store all-gather (run under torchrun on a multi-GPU box):
every rank decodes its own images, stores its rows into every peer's buffer, and after the cross-rank barrier
each rank must hold exactly the rows NCCL's all_gather delivers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import synth
from computervision.pytorch_b200 import ops, distributed as cvd

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, MD = 4, 40
pred = torch.from_numpy(synth.yolov8_pred(100 + rank, B, 8400, nc=80)).to(dev)
det = ops.sort_nms(ops.pred_filter(pred, 80, 0.05), 0.7, max_det=MD, max_nms=30000)
pg = cvd.PeerGather(B, MD, 7, dev)
for i in range(6):                      # alternate the slots, reuse them; the last three through the NVSwitch multicast address
    mc = pg.multicast_ptr(i) if i >= 3 else 0
    if i >= 3 and not mc:
        break
    pg.barrier(i)                       # nobody still reads what the slot held
    pg.buf[(i % pg.depth) * pg.slot_elems:][: pg.slot_elems].fill_(-1.0)
    pg.barrier(i)
    ops.detection_epilogue_allgather(det, ops.ROWS_FULL, pg.peer_ptrs(i), pg.rank, multicast_ptr=mc)
    pg.barrier(i)
    rows, counts = pg.view(i)
    packed = ops.detection_epilogue(det, ops.ROWS_FULL, packed=True)
    ref_rows, ref_counts = cvd.unpack_detections(cvd.gather_packed(packed), B, MD, 7)
    torch.cuda.synchronize()
    assert torch.equal(rows, ref_rows), (rank, i)
    assert torch.equal(counts, ref_counts), (rank, i)
dist.barrier()
if rank == 0:
    print(f"peer gather ok on {world} ranks: {int(ref_counts.sum())} rows identical to NCCL all_gather "
          f"(unicast peer stores{' and multimem.st multicast' if pg.multicast_ptr(0) else '; no multicast address on this box'})")
dist.destroy_process_group()
