"""Lockstep K1 (variant 3) development check: parity vs the oracle on a small case, agreement with
the cluster kernel at bench scale, and timing of both variants (development aid, not the bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from coevonet_b200 import layout, ops
from oracle import weights, mpe_env, rollout as orollout


def pad(rows, in_dim):
    out = np.zeros((rows.shape[0], layout.fc_pitch(in_dim)), dtype=np.float32)
    out[:, :rows.shape[1]] = rows
    return torch.from_numpy(out).cuda()


def small_case(P, K, E, seed, role="agent_0"):
    ms = layout.SEAT_OF[role]
    others = [s for s in range(3) if s != ms]
    seats = layout.SEATS
    counts = {seats[ms]: P, seats[others[0]]: K, seats[others[1]]: K}
    nets = {"adversary_0": weights.make_fc_rows(counts["adversary_0"], 8, seed, 0.02),
            "agent_0": weights.make_fc_rows(counts["agent_0"], 10, seed + 1, 0.02),
            "agent_1": weights.make_fc_rows(counts["agent_1"], 10, seed + 2, 0.02)}
    n = P * K * E
    init = mpe_env.draw_initial_states(n, seed=seed).reshape(P, K, E, 11)
    dev = {r: pad(nets[r], 8 if r == "adversary_0" else 10) for r in nets}
    idx = np.zeros((n, 3), dtype=np.int64)
    e = 0
    for m in range(P):
        for k in range(K):
            for i in range(E):
                for s in range(3):
                    idx[e, s] = m if s == ms else k
                e += 1
    ref = orollout.rollout(nets, idx, init.reshape(-1, 11))
    res = {}
    for variant in (2, 3):
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        out = ops.mpe_rollout(role, dev[seats[ms]], dev[seats[others[0]]], dev[seats[others[1]]],
                              torch.from_numpy(init).cuda(), variant=variant, status=status)
        torch.cuda.synchronize()
        res[variant] = out.cpu().numpy().reshape(n, 4)
        safe = ref["min_gap"] > 1e-4
        err = np.abs(res[variant][:, 0] - ref["sum_good"])
        print(f"{role} P={P} K={K} E={E} variant={variant}: status={int(status.item())} safe={safe.sum()}/{n} "
              f"max|d sum_good| safe={err[safe].max():.3e} all={err.max():.3e} "
              f"max|d gap| safe={np.abs(res[variant][safe, 3] - ref['min_gap'][safe]).max():.3e}", flush=True)


small_case(5, 1, 16, 11)
small_case(40, 1, 16, 12)
small_case(17, 2, 9, 13, role="adversary_0")
small_case(300, 1, 1, 14, role="agent_1")

theta = {"agent_0": pad(weights.make_fc_rows(1, 10, 1), 10),
         "agent_1": pad(weights.make_fc_rows(1, 10, 2), 10),
         "adversary_0": pad(weights.make_fc_rows(1, 8, 3), 8)}
for P, E in ((1024, 16), (4096, 16), (1024, 64)):
    members = ops.es_perturb(theta["agent_0"][0], 10, 0.05, 1, "agent_0", 0, 0, P)
    init = ops.init_states(1, 0, P * E, "cuda").reshape(P, 1, E, 11)
    outs = {}
    for variant in (2, 3):
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
        outs[variant] = out
        print(f"P={P} E={E} variant={variant}: {ms:.3f} ms  {ws/1e6:.2f} M world-steps/s", flush=True)
    a, b = outs[2].reshape(-1, 4), outs[3].reshape(-1, 4)
    safe = torch.minimum(a[:, 3], b[:, 3]) > 1e-4
    same = (a[safe, :3] == b[safe, :3]).all(dim=1)
    print(f"   variants agree bitwise on {int(same.sum())}/{int(safe.sum())} safe episodes "
          f"({a.shape[0]} total); max|d gap| = {float((a[safe, 3] - b[safe, 3]).abs().max()):.3e}", flush=True)
