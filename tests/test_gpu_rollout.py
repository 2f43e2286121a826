"""K1 parity: CUDA rollout kernels (through the C ABI) vs the CPU oracle.

Criterion (BASELINE.json north_star): per-member fitness within 1e-4 relative
(fp32 network, fp64 environment) on identical weights and initial states.  The
policy is an argmax, so an ulp-level difference in fp32 summation order can flip
a decision whose top-2 logit gap is ~0 and fork the trajectory (SURVEY.md
section 7, hard part 2).  Episodes are therefore compared where the oracle's
smallest decision margin exceeds GAP; on those the action traces are identical
and, the environment being bit-exact, the reward sums must agree to 1e-9.  The
remaining episodes are counted and bounded.
"""
import numpy as np
import pytest
import torch

from oracle import layout as olayout
from oracle import mpe_env, rollout as orollout, weights

pytestmark = pytest.mark.gpu

GAP = 1e-4          # decision margin below which a fork is tolerated
RTOL = 1e-4         # north-star tolerance on fitness


def _padded(rows, in_dim):
    from coevonet_b200 import layout
    out = np.zeros((rows.shape[0], layout.fc_pitch(in_dim)), dtype=np.float32)
    out[:, :rows.shape[1]] = rows
    return torch.from_numpy(out).cuda()


def _nets(n_adv, n_a0, n_a1, seed, jitter=0.02):
    return {"adversary_0": weights.make_fc_rows(n_adv, 8, seed, jitter),
            "agent_0": weights.make_fc_rows(n_a0, 10, seed + 1, jitter),
            "agent_1": weights.make_fc_rows(n_a1, 10, seed + 2, jitter)}


def _compare(out, ref, label):
    """out: np[N,4] from the device, ref: oracle dict."""
    safe = ref["min_gap"] > GAP
    n_safe = int(safe.sum())
    assert n_safe >= 0.8 * len(safe), f"{label}: too few safe episodes ({n_safe}/{len(safe)})"
    for col, key in ((0, "sum_good"), (1, "last_good"), (2, "sum_adv")):
        np.testing.assert_allclose(out[safe, col], ref[key][safe], rtol=1e-9, atol=1e-9,
                                   err_msg=f"{label}: {key} differs on safe episodes")
    np.testing.assert_allclose(out[safe, 3], ref["min_gap"][safe], rtol=0, atol=2e-5,
                               err_msg=f"{label}: decision margins differ")
    # forked episodes stay bounded: rewards are distances in a unit box
    rel = np.abs(out[:, 0] - ref["sum_good"]) / np.maximum(1.0, np.abs(ref["sum_good"]))
    n_fork = int((rel > RTOL).sum())
    assert n_fork <= (~safe).sum(), f"{label}: {n_fork} forks but only {(~safe).sum()} unsafe episodes"
    return n_safe, n_fork


def _structured_case(member_role, P, K, E, seed, n_cycles=25, init_shared=False, variant=0, tweak=None):
    from coevonet_b200 import layout, ops
    seats = layout.SEATS
    ms = layout.SEAT_OF[member_role]
    others = [s for s in range(3) if s != ms]
    counts = {seats[ms]: P, seats[others[0]]: K, seats[others[1]]: K}
    nets = _nets(counts["adversary_0"], counts["agent_0"], counts["agent_1"], seed)
    if tweak is not None:
        tweak(nets, ms)
    n = P * K * E
    init_all = mpe_env.draw_initial_states(K * E if init_shared else n, seed=seed)
    init = init_all.reshape((K, E, 11) if init_shared else (P, K, E, 11))
    dev = {r: _padded(nets[r], olayout.OBS_DIM[r]) for r in nets}
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = ops.mpe_rollout(member_role, dev[seats[ms]], dev[seats[others[0]]], dev[seats[others[1]]],
                          torch.from_numpy(init).cuda(), n_cycles=n_cycles, init_shared=init_shared,
                          variant=variant, status=status)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    # oracle on the same episodes
    idx = np.zeros((n, 3), dtype=np.int64)
    flat_init = np.zeros((n, 11))
    e = 0
    for m in range(P):
        for k in range(K):
            for i in range(E):
                for s in range(3):
                    idx[e, s] = m if s == ms else k
                flat_init[e] = init[k, i] if init_shared else init[m, k, i]
                e += 1
    ref = orollout.rollout(nets, idx, flat_init, n_cycles=n_cycles)
    return out.cpu().numpy().reshape(n, 4), ref


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("member_role,P,K,E", [
    ("agent_0", 5, 1, 16), ("agent_1", 3, 2, 8), ("adversary_0", 4, 1, 4),
    ("agent_0", 2, 3, 1), ("adversary_0", 3, 1, 13), ("agent_1", 2, 1, 21),
])
def test_structured_rollout_matches_oracle(variant, member_role, P, K, E):
    out, ref = _structured_case(member_role, P, K, E, seed=1000 + P * 7 + K * 3 + E, variant=variant)
    _compare(out, ref, f"variant{variant} {member_role} P{P} K{K} E{E}")


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_shared_init_and_short_episodes(variant):
    for n_cycles in (0, 1, 8, 24):
        out, ref = _structured_case("agent_0", 3, 2, 5, seed=77, n_cycles=n_cycles,
                                    init_shared=True, variant=variant)
        if n_cycles == 0:
            assert np.all(out[:, :3] == 0.0)
        else:
            _compare(out, ref, f"variant{variant} shared n_cycles={n_cycles}")


def test_cluster_and_generic_kernels_agree_bitwise_on_safe_episodes():
    a, ref = _structured_case("agent_0", 6, 2, 16, seed=5, variant=1)
    b, _ = _structured_case("agent_0", 6, 2, 16, seed=5, variant=2)
    c, _ = _structured_case("agent_0", 6, 2, 16, seed=5, variant=3)
    safe = ref["min_gap"] > GAP
    assert np.array_equal(a[safe, :3], b[safe, :3])
    assert np.array_equal(a[safe, :3], c[safe, :3])


@pytest.mark.parametrize("member_role,P,K,E", [("agent_0", 40, 1, 16), ("adversary_0", 17, 2, 9),
                                               ("agent_1", 300, 1, 1)])
def test_lockstep_kernels_match_oracle_on_multi_tile_shapes(member_role, P, K, E):
    """Variant 3 (tcgen05 opponents): more than one 128-episode tile per opponent, a ragged last
    tile, ragged 16-episode member chunks; 3xTF32 keeps the opponents' logits at fp32 accuracy, so
    the same margin criterion as the fp32 kernels applies."""
    out, ref = _structured_case(member_role, P, K, E, seed=4242 + P, variant=3)
    _compare(out, ref, f"lockstep {member_role} P{P} K{K} E{E}")


def test_lockstep_is_the_auto_choice_for_es_shapes_and_counts_its_launches():
    from coevonet_b200 import ops
    # opponent max + split + layer-1 statistics, member layer-1 statistics, initial states, 3 kernels per world step
    assert ops.rollout_plan(0, 1024, 1, 16) == (3, 5 + 3 * 25)
    assert ops.rollout_plan(0, 20, 3, 1)[0] == 2
    assert ops.rollout_plan(0, 8192, 1, 1)[0] == 3          # GA shape at scale: lockstep too
    assert ops.rollout_plan(0, 1, 1, 10)[0] == 2
    assert ops.rollout_plan(0, 1024, 1, 16, variant=2) == (2, 3)


def test_indexed_rollout_matches_golden_reference_episodes(golden):
    """The committed outputs of the reference's own play_game (episodes.npz)."""
    from coevonet_b200 import ops
    g = golden("episodes")
    seeds = g["seeds"]
    nt = int(g["n_triples"])
    jit = float(g["ln_jitter"])
    nets = {"agent_0": weights.make_fc_rows(nt, 10, int(seeds[0]), jit),
            "agent_1": weights.make_fc_rows(nt, 10, int(seeds[1]), jit),
            "adversary_0": weights.make_fc_rows(nt, 8, int(seeds[2]), jit)}
    idx = torch.from_numpy(g["idx"].astype(np.int32)).cuda()
    init = torch.from_numpy(g["init"]).cuda()
    out = ops.mpe_rollout_indexed(_padded(nets["adversary_0"], 8), _padded(nets["agent_0"], 10),
                                  _padded(nets["agent_1"], 10), idx, init,
                                  pos_first=bool(g["integrate_pos_first"]))
    s0, s1, sadv = ops.reward_slots(out)           # rotated attribution, Appendix B
    got = torch.stack([s0, s1, sadv], dim=1).cpu().numpy()
    safe = g["min_gap"] > GAP
    assert safe.sum() >= 0.8 * len(safe)
    np.testing.assert_allclose(got[safe], g["rewards"][safe], rtol=RTOL, atol=1e-9)


def test_step_limit_slots_match_golden(golden):
    from coevonet_b200 import ops
    g = golden("step_limit")
    seeds = g["seeds"]
    nets = {"agent_0": weights.make_fc_rows(1, 10, int(seeds[0])),
            "agent_1": weights.make_fc_rows(1, 10, int(seeds[1])),
            "adversary_0": weights.make_fc_rows(1, 8, int(seeds[2]))}
    dev = {r: _padded(nets[r], olayout.OBS_DIM[r]) for r in nets}
    for L, init, want in zip(g["limits"], g["init"], g["rewards"]):
        nc = ops.cycles_for_limit(int(L))
        out = ops.mpe_rollout("agent_0", dev["agent_0"], dev["adversary_0"], dev["agent_1"],
                              torch.from_numpy(init.reshape(1, 1, 1, 11)).cuda(), n_cycles=nc)
        s0, s1, sadv = ops.reward_slots(out.reshape(1, 4), agent_step_limit=int(L))
        got = np.array([float(s0), float(s1), float(sadv)])
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-12, err_msg=f"L={L}")


def test_nonfinite_weights_raise_like_the_reference():
    from coevonet_b200 import ops
    nets = _nets(1, 2, 1, 9)
    nets["agent_0"][1, 100] = np.nan
    dev = {r: _padded(nets[r], olayout.OBS_DIM[r]) for r in nets}
    init = torch.from_numpy(mpe_env.draw_initial_states(2 * 4).reshape(2, 1, 4, 11)).cuda()
    for variant in (1, 2, 3):
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        ops.mpe_rollout("agent_0", dev["agent_0"], dev["adversary_0"], dev["agent_1"], init,
                        variant=variant, status=status)
        with pytest.raises(ValueError):
            ops.raise_on_status(status)


def test_full_size_properties():
    """BASELINE config-2 shape (P=1024 x 16 envs): determinism, member
    permutation equivariance and agreement with the oracle on a sample."""
    from coevonet_b200 import layout, ops
    P, E = 1024, 16
    gen = torch.Generator(device="cpu").manual_seed(3)
    base = _padded(weights.make_fc_rows(1, 10, 31), 10)
    members = base.repeat(P, 1)
    members[:, :layout.fc_dim(10)] += 0.05 * torch.randn((P, layout.fc_dim(10)), generator=gen).cuda()
    a1 = _padded(weights.make_fc_rows(1, 10, 32), 10)
    adv = _padded(weights.make_fc_rows(1, 8, 33), 8)
    init = ops.init_states(1870300, 0, P * E, "cuda").reshape(P, 1, E, 11)
    out1 = ops.mpe_rollout("agent_0", members, adv, a1, init)
    out2 = ops.mpe_rollout("agent_0", members, adv, a1, init)
    assert torch.equal(out1, out2), "rollout is not deterministic"
    perm = torch.randperm(P, generator=gen).cuda()
    out3 = ops.mpe_rollout("agent_0", members[perm].contiguous(), adv, a1, init[perm].contiguous())
    assert torch.equal(out3, out1[perm]), "rollout depends on member order"
    # oracle on 6 members
    sel = [0, 1, 511, 512, 1022, 1023]
    nets = {"agent_0": members[sel].cpu().numpy(), "agent_1": a1.cpu().numpy(),
            "adversary_0": adv.cpu().numpy()}
    idx = np.array([[0, i, 0] for i in range(len(sel)) for _ in range(E)])
    flat = init[sel].cpu().numpy().reshape(-1, 11)
    ref = orollout.rollout(nets, idx, flat)
    _compare(out1[sel].cpu().numpy().reshape(-1, 4), ref, "full-size sample")
    # rewards are bounded by geometry: |r| <= 25 * (2*sqrt(2) + drift)
    assert torch.isfinite(out1).all()
    assert (out1[..., 2] <= 0).all()


def test_large_population_stress_is_deterministic():
    """BASELINE config-3 scale per GPU (tens of thousands of members, hof = 3): two runs of
    ~150 k tiles must complete and agree bit for bit.  (Regression: a parity-aliased mbarrier
    wait in the streamed-operand pipeline only misfired at this scale.)"""
    from coevonet_b200 import ops
    P, K = 49152, 3
    pop = ops.fc_init(10, 7, "agent_0", 0, P, "cuda")
    adv = ops.fc_init(8, 7, "adversary_0", P, K, "cuda")
    a1 = ops.fc_init(10, 7, "agent_1", P, K, "cuda")
    init = ops.init_states(7, 0, P * K, "cuda").reshape(P, K, 1, 11)
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    out1 = ops.mpe_rollout("agent_0", pop, adv, a1, init, status=status)
    out2 = ops.mpe_rollout("agent_0", pop, adv, a1, init, status=status)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert torch.equal(out1, out2)
    # spot-check 4 members against the oracle
    sel = [0, 12345, 30000, P - 1]
    nets = {"agent_0": pop[sel].cpu().numpy(), "agent_1": a1.cpu().numpy(), "adversary_0": adv.cpu().numpy()}
    idx = np.array([[k, i, k] for i in range(len(sel)) for k in range(K)])
    ref = orollout.rollout(nets, idx, init[sel].cpu().numpy().reshape(-1, 11))
    _compare(out1[sel].cpu().numpy().reshape(-1, 4), ref, "large-population sample")


def test_lockstep_cta_pair_form_matches_oracle():
    """The opt-in CTA-pair form of the opponent kernel (tcgen05 cta_group::2, M = 256 per MMA, B halves
    exchanged over the pair link; CEV_LS_PAIR=1 is read once per process, so it runs in a child)."""
    import os, subprocess, sys
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import test_gpu_rollout as t\n"
        "out, ref = t._structured_case('agent_1', 70, 2, 9, seed=99, variant=3)\n"
        "print(t._compare(out, ref, 'pair form'))\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, CEV_LS_PAIR="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("P,K,E", [(192, 1, 16), (40, 2, 9)])
def test_roles_in_one_pass_equal_one_pass_per_role(P, K, E):
    """cev_mpe_rollout_roles_f32 (the three roles' member / opponent kernels interleaved on three
    streams) returns bit for bit what three cev_mpe_rollout_f32 calls with the lockstep kernels do."""
    from coevonet_b200 import layout, ops
    nets = _nets(max(P, K), max(P, K), max(P, K), seed=777)
    rows = {r: _padded(nets[r], layout.OBS_DIM[r]) for r in nets}
    specs = []
    for i, role in enumerate(("agent_0", "agent_1", "adversary_0")):
        ms = layout.SEAT_OF[role]
        others = [layout.SEATS[s] for s in range(3) if s != ms]
        init = ops.init_states(99, i, P * K * E, "cuda").reshape(P, K, E, 11)
        specs.append((role, rows[role][:P].contiguous(), rows[others[0]][:K].contiguous(),
                      rows[others[1]][:K].contiguous(), init))
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    fused = ops.mpe_rollout_roles(specs, variant=3, status=status)
    torch.cuda.synchronize()
    for spec, out in zip(specs, fused):
        single = ops.mpe_rollout(spec[0], spec[1], spec[2], spec[3], spec[4], variant=3, status=status)
        assert torch.equal(single, out), f"{spec[0]}: fused pass differs from the single-role pass"
    assert int(status.item()) == 0


@pytest.mark.parametrize("w2_scale,g1_scale", [(250.0, 1.0), (0.01, 1.0), (1.0, 40.0), (30.0, 0.02)])
def test_lockstep_opponents_with_extreme_weight_scales(w2_scale, g1_scale):
    """The opponent kernel scales its FP16 operands by powers of two taken from max |fc2.W| and from the LayerNorm-1
    parameters: opponents whose fc2 matrix or ln1.gamma are orders of magnitude away from an initialised network
    must still agree with the fp32 oracle (same margin criterion as every other rollout test)."""
    def tweak(nets, ms):
        for seat, role in enumerate(("adversary_0", "agent_0", "agent_1")):
            if seat == ms:
                continue
            in_dim = olayout.OBS_DIM[role]
            rows = nets[role]
            fc1 = 512 * in_dim
            rows[:, fc1 + 512:fc1 + 1024] *= np.float32(g1_scale)                       # ln1.gamma
            fc2w = fc1 + 3 * 512
            rows[:, fc2w:fc2w + 256 * 512] *= np.float32(w2_scale)                       # fc2.W
            # keep the pre-LayerNorm-2 scale sane for the comparison: LayerNorm-2 normalises it away anyway
    out, ref = _structured_case("agent_0", 24, 1, 16, seed=4321, variant=3, tweak=tweak)
    _compare(out, ref, f"lockstep opponents fc2.W x{w2_scale} ln1.gamma x{g1_scale}")
