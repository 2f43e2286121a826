"""Stress of the roles-in-one-pass path: large and ragged shapes, fused == one pass per role bit for bit, repeated
runs identical (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coevonet_b200 import layout, ops

theta = {"agent_0": ops.fc_init(10, 1, "agent_0", 0, 3, "cuda"), "agent_1": ops.fc_init(10, 2, "agent_1", 0, 3, "cuda"),
         "adversary_0": ops.fc_init(8, 3, "adversary_0", 0, 3, "cuda")}
for P, K, E in ((4096, 1, 16), (8192, 1, 1), (1000, 3, 7), (149, 2, 33), (2048, 1, 17)):
    specs = []
    for i, role in enumerate(("agent_0", "agent_1", "adversary_0")):
        ms = layout.SEAT_OF[role]
        others = [layout.SEATS[s] for s in range(3) if s != ms]
        members = ops.es_perturb(theta[role][0], layout.OBS_DIM[role], 0.05, 7, role, 0, 0, P)
        init = ops.init_states(3, i, P * K * E, "cuda").reshape(P, K, E, 11)
        specs.append((role, members, theta[others[0]][:K].contiguous(), theta[others[1]][:K].contiguous(), init))
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    ref = None
    for rep in range(3):
        fused = ops.mpe_rollout_roles(specs, variant=3, status=status)
        torch.cuda.synchronize()
        if ref is None:
            ref = [f.clone() for f in fused]
            single = [ops.mpe_rollout(s[0], s[1], s[2], s[3], s[4], variant=3, status=status) for s in specs]
            assert all(torch.equal(a, b) for a, b in zip(fused, single)), f"fused != single at {(P, K, E)}"
        else:
            assert all(torch.equal(a, b) for a, b in zip(fused, ref)), f"run {rep} differs at {(P, K, E)}"
    assert int(status.item()) == 0
    print(f"P={P} K={K} E={E}: fused == single, 3 runs identical, mean reward {float(ref[0][..., 0].mean()):.4f}", flush=True)
