"""Timing of the lockstep rollout (variant 3) at bench shape; CEV_LS_SKIP=1/2 isolates the member /
opponent kernel (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coevonet_b200 import layout, ops


theta = {"agent_0": ops.fc_init(10, 1, "agent_0", 0, 1, "cuda"), "agent_1": ops.fc_init(10, 2, "agent_1", 0, 1, "cuda"),
         "adversary_0": ops.fc_init(8, 3, "adversary_0", 0, 1, "cuda")}
shapes = [(1024, 16)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for P, E in shapes:
    members = ops.es_perturb(theta["agent_0"][0], 10, 0.05, 1, "agent_0", 0, 0, P)
    init = ops.init_states(1, 0, P * E, "cuda").reshape(P, 1, E, 11)
    for _ in range(2):
        out = ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 5
    for _ in range(n):
        out = ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=3)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"skip={os.environ.get('CEV_LS_SKIP','0')} P={P} E={E}: {ms:.3f} ms per rollout, {ms/25*1000:.1f} us per cycle, "
          f"{P*E*25/ms/1e3:.2f} M world-steps/s", flush=True)
