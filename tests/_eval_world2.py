"""Helper of tests/test_eval_gpu.py (run under torchrun, 2 ranks, gloo, one GPU): every rank evaluates its shard of
7 YOLOv8 images + 5 YOLOv7 images and the merged lists must equal the single-process result."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle  # noqa: E402
import synth  # noqa: E402
from computervision.pytorch_b200.core.eval import BatchedDetectionEvaluator, YOLOv7Evaluator  # noqa: E402

DEV = "cuda:0"
heads = [torch.from_numpy(l).to(DEV) for l in synth.yolov8_head(78, B=7)]
v7 = [torch.from_numpy(l).to(DEV) for l in synth.yolov7_head(79, B=5)]
hw = [(480, 640), (333, 500), (640, 427), (1080, 1920), (640, 640), (375, 500), (500, 375)]
names = [f"c{i}" for i in range(80)]
catid = list(range(1, 81))


def run():
    ev8 = BatchedDetectionEvaluator(80, synth.YOLOV8_STRIDES, (640, 640), batch_size=3)
    ev7 = YOLOv7Evaluator(80, oracle.YOLOV7_ANCHORS, oracle.YOLOV7_MASK, (640, 640), batch_size=2, max_out=8192)
    f8 = lambda idx: ([l[idx] for l in heads], [hw[i] for i in idx])    # noqa: E731
    f7 = lambda idx: ([l[idx] for l in v7], [hw[i] for i in idx])       # noqa: E731
    return (ev8.evaluate_coco(7, f8, list(range(100, 107)), catid), ev8.evaluate_voc(7, f8, names),
            ev7.evaluate_voc(5, f7, names), ev7.evaluate_coco(5, f7, list(range(5)), catid))


single = run()                      # before init_process_group: the whole list on every rank
dist.init_process_group("gloo")
assert dist.get_world_size() == 2
sharded = run()
assert list(BatchedDetectionEvaluator.my_indices(7)) == ([0, 1, 2, 3] if dist.get_rank() == 0 else [4, 5, 6])
for a, b in zip(single, sharded):
    assert a == b
assert len(single[0]) > 1000 and len(single[1]) == 7 and len(single[2]) == 5
dist.barrier()
if dist.get_rank() == 0:
    print("world2 ok")
dist.destroy_process_group()
