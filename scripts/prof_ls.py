"""One lockstep rollout at bench shape for ncu (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coevonet_b200 import layout, ops
from oracle import weights

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1

def pad(rows, in_dim):
    out = np.zeros((rows.shape[0], layout.fc_pitch(in_dim)), dtype=np.float32)
    out[:, :rows.shape[1]] = rows
    return torch.from_numpy(out).cuda()

theta = {"agent_0": pad(weights.make_fc_rows(1, 10, 1), 10),
         "agent_1": pad(weights.make_fc_rows(1, 10, 2), 10),
         "adversary_0": pad(weights.make_fc_rows(1, 8, 3), 8)}
members = ops.es_perturb(theta["agent_0"][0], 10, 0.05, 1, "agent_0", 0, 0, P)
init = ops.init_states(1, 0, P * E, "cuda").reshape(P, 1, E, 11)
for _ in range(reps):
    out = ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=3)
torch.cuda.synchronize()
print("ok", float(out[..., 0].mean()))
