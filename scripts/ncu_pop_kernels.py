"""The population kernels (K3, K5, K6 both forms, K7, weight statistics) a few times each at bench scale, for
`ncu -k regex:...` captures and launch lists (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coevonet_b200 import layout, ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
in_dim = 10
D, pitch = layout.fc_dim(in_dim), layout.fc_pitch(in_dim)
theta = ops.fc_init(in_dim, 1, "agent_0", 0, 1, dev)[0].contiguous()
elites = ops.fc_init(in_dim, 1, "agent_0", 1, 5, dev)
rows = torch.empty((P, pitch), dtype=torch.float32, device=dev)
fit = torch.linspace(-30, -5, P, dtype=torch.float64, device=dev)
for g in range(reps):
    ops.es_perturb(theta, in_dim, 0.05, 1, "agent_0", g, 0, P, out=rows)
    ops.es_update_members(fit, rows, theta, in_dim, 0.05, 0.1, P)
    ops.es_update(fit, in_dim, 0.05, 0.1, P, 1, "agent_0", g, 0)
    ops.diversity_dist(rows, theta, in_dim)
    ops.weight_stats(rows, in_dim)
    ops.ga_repopulate(elites, D, 0.05, 1, "agent_0", g, 0, P, out=rows)
    ops.select_topk(fit, 5)
torch.cuda.synchronize()
print("ok")
