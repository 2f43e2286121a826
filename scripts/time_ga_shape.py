"""GA-shaped rollouts (one or three episodes per member): cluster form vs lockstep form (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coevonet_b200 import ops
for P, K in ((8192, 1), (8192, 3), (32768, 1)):
    pop = ops.fc_init(10, 7, "agent_0", 0, P, "cuda")
    adv = ops.fc_init(8, 7, "adversary_0", P, K, "cuda")
    a1 = ops.fc_init(10, 7, "agent_1", P, K, "cuda")
    init = ops.init_states(7, 0, P * K, "cuda").reshape(P, K, 1, 11)
    outs = {}
    for variant in (2, 3):
        for _ in range(1):
            out = ops.mpe_rollout("agent_0", pop, adv, a1, init, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = ops.mpe_rollout("agent_0", pop, adv, a1, init, variant=variant)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        outs[variant] = out
        print(f"P={P} K={K} E=1 variant={variant}: {ms:.1f} ms  {P*K*25/ms/1e3:.2f} M world-steps/s", flush=True)
    a, b = outs[2].reshape(-1, 4), outs[3].reshape(-1, 4)
    safe = torch.minimum(a[:, 3], b[:, 3]) > 1e-4
    same = (a[safe, :3] == b[safe, :3]).all(dim=1)
    print(f"   agree bitwise on {int(same.sum())}/{int(safe.sum())} safe episodes ({a.shape[0]} total)", flush=True)
