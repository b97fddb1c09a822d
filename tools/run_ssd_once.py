"""Runs the SSD decode+filter kernel a few times on the C4-sized synthetic workload (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import bench_paths as bp
bp.bench_ssd(3)
