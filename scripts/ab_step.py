"""A/B timing of ESEngine.step() variants on one box (development aid)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from coevonet_b200 import engine, layout
from coevonet_b200.MPE.fcnetwork import FCNetwork

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
args = bench._args_bag(bench.P_PER_GPU)
torch.manual_seed(0)
theta = {r: FCNetwork(layout.OBS_DIM[r], 5, "float32").flat_row() for r in bench.ROLES}
eng = engine.ESEngine(args, dev, theta)
def run(n=8):
    for _ in range(2): eng.step(sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.step(sync=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(2):
    for ov in (False, True):
        eng.overlap_roles = ov
        print(f"overlap_roles={ov}: {run():.3f} ms per step", flush=True)
