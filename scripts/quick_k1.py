"""Quick K1 timing sweep (development aid, not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from coevonet_b200 import layout, ops, _lib


print("device_info (n_sm, n_clusters):", _lib.device_info(0))
print("fp32 peak FFMA  :", ops.fp32_peak("cuda:0", 0))
print("fp32 peak FFMA2 :", ops.fp32_peak("cuda:0", 1))
theta = {"agent_0": ops.fc_init(10, 1, "agent_0", 0, 1, "cuda"), "agent_1": ops.fc_init(10, 2, "agent_1", 0, 1, "cuda"),
         "adversary_0": ops.fc_init(8, 3, "adversary_0", 0, 1, "cuda")}
for P, E in ((1024, 16), (1024, 8), (1024, 4), (4096, 16)):
    members = ops.es_perturb(theta["agent_0"][0], 10, 0.05, 1, "agent_0", 0, 0, P)
    init = ops.init_states(1, 0, P * E, "cuda").reshape(P, 1, E, 11)
    for variant in (2, 1):
        if variant == 1 and P * E > 20000: continue
        for _ in range(2):
            out = ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 3
        for _ in range(n):
            out = ops.mpe_rollout("agent_0", members, theta["adversary_0"], theta["agent_1"], init, variant=variant)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        ws = P * E * 25 / (ms * 1e-3)
        print(f"P={P} E={E} variant={variant}: {ms:.3f} ms  {ws/1e6:.2f} M world-steps/s  "
              f"{ws*822784/1e12:.2f} TFLOP/s")
