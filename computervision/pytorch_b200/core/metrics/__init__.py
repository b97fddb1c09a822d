"""Evaluation metrics of the detection path."""
from .mAP import get_map_from_rows, voc_ap  # noqa: F401
