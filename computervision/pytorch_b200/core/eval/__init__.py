"""Batched evaluation driver + detection serialisation: the steps that follow NMS in
the reference's `evaluate_on_voc` / `evaluate_on_coco` loops, batched and sharded by image."""
from .serialize import coco_results, voc_lines  # noqa: F401
from .driver import (BatchedDetectionEvaluator, CenterNetEvaluator, SsdEvaluator,  # noqa: F401
                     YOLOv7Evaluator)
