"""Timing of the three roles' lockstep rollouts of one generation in one pass (cev_mpe_rollout_roles_f32) at bench
shape, next to one pass per role; also the host time the launches take (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from coevonet_b200 import layout, ops

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
E = int(sys.argv[2]) if len(sys.argv) > 2 else 16
theta = {"agent_0": ops.fc_init(10, 1, "agent_0", 0, 1, "cuda"), "agent_1": ops.fc_init(10, 2, "agent_1", 0, 1, "cuda"),
         "adversary_0": ops.fc_init(8, 3, "adversary_0", 0, 1, "cuda")}
specs = []
for i, role in enumerate(("agent_0", "agent_1", "adversary_0")):
    ms = layout.SEAT_OF[role]
    others = [layout.SEATS[s] for s in range(3) if s != ms]
    members = ops.es_perturb(theta[role][0], layout.OBS_DIM[role], 0.05, 1, role, 0, 0, P)
    init = ops.init_states(1, i, P * E, "cuda").reshape(P, 1, E, 11)
    specs.append((role, members, theta[others[0]], theta[others[1]], init))


def run(fused):
    if fused:
        return ops.mpe_rollout_roles(specs, variant=3)
    return [ops.mpe_rollout(s[0], s[1], s[2], s[3], s[4], variant=3) for s in specs]


for fused in (True, False):
    for _ in range(2):
        run(fused)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        run(fused)
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"fused={fused} P={P} E={E}: {ms:.3f} ms per 3-role pass ({ms / 75 * 1000:.1f} us per role-step), "
          f"host enqueue {host:.3f} ms, {3 * P * E * 25 / ms / 1e3:.2f} M world-steps/s", flush=True)
