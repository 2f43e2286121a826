"""One lockstep rollout at bench shape for ncu (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coevonet_b200 import layout, ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1


theta = {"agent_0": ops.fc_init(10, 1, "agent_0", 0, 1, "cuda"), "agent_1": ops.fc_init(10, 2, "agent_1", 0, 1, "cuda"),
         "adversary_0": ops.fc_init(8, 3, "adversary_0", 0, 1, "cuda")}
members = ops.es_perturb(theta["agent_0"][0], 10, 0.05, 1, "agent_0", 0, 0, P)
init = ops.init_states(1, 0, P * E, "cuda").reshape(P, 1, E, 11)
for _ in range(reps):
    out = ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=3)
torch.cuda.synchronize()
print("ok", float(out[..., 0].mean()))
