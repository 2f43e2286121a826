"""Phases of an ES generation on the main stream, from CUDA events recorded INSIDE consecutive real steps (development aid):
perturb (joined), rollout pass + reward slots, update, generation end."""
import os, sys
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
marks = []
orig_run_roles, orig_rollouts, orig_update, orig_finish = eng._run_roles, eng._rollouts, eng.update, eng._finish_generation


def ev(tag):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((tag, e))


def run_roles(fn):
    ev("perturb0"); orig_run_roles(fn); ev("perturb1")
def rollouts(specs, limit):
    ev("roll0"); out = orig_rollouts(specs, limit); ev("roll1"); return out
def update():
    ev("upd0"); orig_update(); ev("upd1")
def finish(*a):
    ev("fin0"); orig_finish(*a); ev("fin1")
eng._run_roles, eng._rollouts, eng.update, eng._finish_generation = run_roles, rollouts, update, finish
for _ in range(3):
    eng.step(sync=False)
torch.cuda.synchronize()
marks.clear()
n = 8
for _ in range(n):
    eng.step(sync=False)
ev("end")
torch.cuda.synchronize()
tot = {}
for (t0, e0), (t1, e1) in zip(marks[:-1], marks[1:]):
    tot[f"{t0}->{t1}"] = tot.get(f"{t0}->{t1}", 0.0) + e0.elapsed_time(e1)
for k, v in tot.items():
    print(f"{k:22s} {v / n:8.3f} ms")
print(f"{'sum':22s} {sum(tot.values()) / n:8.3f} ms per generation")
