"""Tiny launches (evaluation games, train_GA.sh-sized populations): generic vs cluster K1 (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coevonet_b200 import ops
for P, K, E in ((1, 1, 10), (20, 1, 1), (20, 3, 1), (64, 1, 1)):
    pop = ops.fc_init(10, 7, "agent_0", 0, P, "cuda")
    adv = ops.fc_init(8, 7, "adversary_0", P, K, "cuda")
    a1 = ops.fc_init(10, 7, "agent_1", P, K, "cuda")
    init = ops.init_states(7, 0, P * K * E, "cuda").reshape(P, K, E, 11)
    res = []
    for variant in (1, 2):
        for _ in range(2):
            out = ops.mpe_rollout("agent_0", pop, adv, a1, init, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = ops.mpe_rollout("agent_0", pop, adv, a1, init, variant=variant)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 5 * 1e3)
    print(f"P={P} K={K} E={E}: generic {res[0]:.0f} us, cluster {res[1]:.0f} us", flush=True)
