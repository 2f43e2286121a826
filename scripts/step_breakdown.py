"""Where one ES generation spends its device time: perturb, rollouts, update, evaluation games, each timed alone with
CUDA events in a steady loop, next to the whole step (development aid)."""
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


def timed(fn, n=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


eng.step(sync=False)
print(f"step            {timed(lambda: eng.step(sync=False)):.3f} ms")
print(f"evaluate        {timed(eng.evaluate):.3f} ms   (perturb + rollouts + reward slots)")
print(f"update          {timed(eng.update):.3f} ms")
t = eng.theta
def fin():
    eng._finish_generation(t["agent_0"], t["agent_1"], t["adversary_0"]); eng._join_eval()
print(f"eval games      {timed(fin):.3f} ms   (alone, joined)")
def perturb():
    for r in bench.ROLES:
        eng.k.es_perturb(eng.theta[r], layout.OBS_DIM[r], eng.sigma_dev(r), eng.seed, r, eng.gen, eng.shard.row0,
                         eng.shard.n_local, out=eng.members[r])
print(f"perturb x3      {timed(perturb):.3f} ms")
def ev_noeval():
    eng.evaluate(); eng.update()
print(f"evaluate+update {timed(ev_noeval):.3f} ms")
