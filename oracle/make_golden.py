"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Run in the build container
only (``/root/reference`` does not exist on the GPU box):

    python -m oracle.make_golden

Every fixture stores the inputs as seeds (weights come from
``oracle.weights`` = NumPy PCG64, reproducible anywhere) plus the reference's
outputs, so the files stay small.  What produced each output is the reference
function named in the fixture's ``source`` field.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from . import ga_es, layout, mpe_env, philox, stubs, weights
from . import deepqn as odqn
from . import rollout as orollout

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                          "tests", "golden")
IN_DIM = layout.OBS_DIM


def _load_fc(ref, env, args, role, row):
    """A reference MPEAgent whose FCNetwork holds ``row``."""
    agent = ref.utils_game_logic_functions.create_agent(env, args, role)
    segs, _ = layout.fc_segments(IN_DIM[role])
    sd = {name: torch.from_numpy(np.array(row[off:off + int(np.prod(shape))]).reshape(shape))
          for name, off, shape, _ in segs}
    agent.model.load_state_dict(sd)
    return agent


def golden_fc_logits(ref, args):
    """FCNetwork.forward (MPE/fcnetwork.py:37-70) on fixed weights / obs."""
    rng = np.random.Generator(np.random.PCG64(11))
    out = {}
    env = ref.utils_game_logic_functions.initialize_env(args)
    for role, seed in (("agent_0", 101), ("adversary_0", 102)):
        rows = weights.make_fc_rows(3, IN_DIM[role], seed, ln_jitter=0.05)
        obs = rng.uniform(-2, 2, (16, IN_DIM[role])).astype(np.float32)
        logits = np.zeros((3, 16, 5), dtype=np.float32)
        acts = np.zeros((3, 16), dtype=np.int32)
        for m in range(3):
            agent = _load_fc(ref, env, args, role, rows[m])
            for b in range(16):
                with torch.no_grad():
                    logits[m, b] = agent.model.forward(torch.from_numpy(obs[b]), args).numpy()
                    acts[m, b] = agent.model.determine_action(torch.from_numpy(obs[b]), args)
        mine = np.stack([orollout.fc_forward(rows[m], obs, IN_DIM[role]) for m in range(3)])
        print(f"[fc_logits {role}] max |oracle-ref| = {np.abs(mine - logits).max():.3e}")
        out[f"{role}.seed"] = seed
        out[f"{role}.obs"] = obs
        out[f"{role}.logits"] = logits
        out[f"{role}.actions"] = acts
    np.savez_compressed(os.path.join(GOLDEN_DIR, "fc_logits.npz"),
                        source="MPE/fcnetwork.py:37-90 FCNetwork.forward/determine_action",
                        ln_jitter=0.05, **out)


def golden_episodes(ref, args, n_triples=6, games_per_triple=8):
    """play_game (utils/game_logic_functions.py:215) on the restated env:
    returned reward triple + action trace + initial state per game."""
    glf = ref.utils_game_logic_functions
    env = glf.initialize_env(args)           # reset(seed=1870300)
    rows = {r: weights.make_fc_rows(n_triples, IN_DIM[r], s, ln_jitter=0.02)
            for r, s in (("agent_0", 201), ("agent_1", 202), ("adversary_0", 203))}
    n = n_triples * games_per_triple
    ref_rewards = np.zeros((n, 3))
    idx = np.zeros((n, 3), dtype=np.int64)
    traces = np.zeros((n, 25, 3), dtype=np.int32)
    e = 0
    for t in range(n_triples):
        agents = {r: _load_fc(ref, env, args, r, rows[r][t]) for r in rows}
        # record actions by wrapping determine_action
        rec = []
        for r in rows:
            m = agents[r].model
            orig = m.determine_action
            m.determine_action = (lambda o, a, _orig=orig: rec.append(_orig(o, a)) or rec[-1])
        for g in range(games_per_triple):
            rec.clear()
            r0, r1, radv = glf.play_game(env, agents["agent_0"].model, agents["agent_1"].model,
                                         agents["adversary_0"].model, args)
            ref_rewards[e] = (r0, r1, radv)
            traces[e] = np.asarray(rec, dtype=np.int32).reshape(25, 3)
            idx[e] = (t, t, t)
            e += 1
    init = np.stack(env.init_state_log[1:1 + n])    # log[0] is initialize_env's own reset
    res = orollout.rollout({"adversary_0": rows["adversary_0"], "agent_0": rows["agent_0"],
                            "agent_1": rows["agent_1"]}, idx, init,
                           pos_first=mpe_env.INTEGRATE_POS_FIRST)
    s0, s1, sadv = orollout.compat_slots(res)
    mine = np.stack([s0, s1, sadv], axis=1)
    same = (res["actions"] == traces).all(axis=(1, 2))
    print(f"[episodes] action traces identical: {same.sum()}/{n}; "
          f"max |reward diff| on identical traces = {np.abs(mine - ref_rewards)[same].max():.3e}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "episodes.npz"),
                        source="utils/game_logic_functions.py:215 play_game on oracle.mpe_env",
                        integrate_pos_first=mpe_env.INTEGRATE_POS_FIRST,
                        seeds=np.array([201, 202, 203]), ln_jitter=0.02,
                        n_triples=n_triples, idx=idx, init=init, rewards=ref_rewards,
                        actions=traces, min_gap=res["min_gap"])


def golden_step_limit(ref):
    """play_MPE under agent-step limits L < 75 (game_logic_functions.py:127,195)."""
    glf = ref.utils_game_logic_functions
    rows = {r: weights.make_fc_rows(1, IN_DIM[r], s)
            for r, s in (("agent_0", 301), ("agent_1", 302), ("adversary_0", 303))}
    limits = np.array([1, 2, 3, 4, 5, 6, 30, 73, 74, 75, 400])
    out = np.zeros((len(limits), 3))
    inits = []
    for i, L in enumerate(limits):
        args = stubs.RefArgs(max_timesteps_per_episode=int(L))
        env = glf.initialize_env(args)
        agents = {r: _load_fc(ref, env, args, r, rows[r][0]) for r in rows}
        out[i] = glf.play_game(env, agents["agent_0"].model, agents["agent_1"].model,
                               agents["adversary_0"].model, args)
        inits.append(env.init_state_log[1])
        nc = orollout.cycles_for_limit(int(L))
        res = orollout.rollout({k: v for k, v in rows.items()}, np.zeros((1, 3), dtype=int),
                               inits[-1][None], n_cycles=nc)
        mine = np.array([s[0] for s in orollout.compat_slots(res, int(L))])
        print(f"[step_limit L={L}] ref={out[i]} oracle={mine}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "step_limit.npz"),
                        source="utils/game_logic_functions.py:123-212 play_MPE with limits",
                        seeds=np.array([301, 302, 303]), limits=limits,
                        init=np.stack(inits), rewards=out)


class _NoiseFeeder:
    """Replaces torch.normal / np.random.normal inside the reference so its
    mutation arithmetic runs on noise we control."""

    def __init__(self, z):
        self.z = np.asarray(z, dtype=np.float32).reshape(-1)
        self.pos = 0

    def torch_normal(self, mean, std, size):
        n = int(np.prod(size))
        out = self.z[self.pos:self.pos + n].reshape(tuple(size))
        self.pos += n
        # sigma * z in fp32, like the device
        return torch.from_numpy((np.float32(std) * out).astype(np.float32))

    def np_normal(self, loc=0.0, scale=1.0, size=None):
        n = int(np.prod(size))
        out = self.z[self.pos:self.pos + n]
        self.pos += n
        # fp32 product widened: agent.py:52 adds fp64 noise to fp32 weights
        return (np.float32(scale) * out).astype(np.float32).astype(np.float64)


def golden_mutation(ref, args):
    """Agent.mutate (agent.py:25-29) and Agent.mutate_ES (agent.py:31-70)
    with injected Philox noise; compute_weight_update; diversity_penalty."""
    glf = ref.utils_game_logic_functions
    env = glf.initialize_env(args)
    role = "agent_0"
    D = layout.fc_dim(10)
    pidx = layout.fc_perturbable_index(10)
    parent = weights.make_fc_rows(1, 10, 401, ln_jitter=0.02)[0]
    # --- GA ---
    z = philox.normals(1870300, philox.KIND_GA, philox.ROLE_ID[role], 3, [7], D)[0]
    agent = _load_fc(ref, env, args, role, parent)
    feeder = _NoiseFeeder(z)
    orig = torch.normal
    torch.normal = feeder.torch_normal
    try:
        agent.mutate(0.05)
    finally:
        torch.normal = orig
    child = layout.pack_fc_state_dict(agent.model.state_dict(), 10)
    mine = ga_es.ga_repopulate(np.stack([parent, parent]), [0], 0.05, np.stack([z, z]))[1]
    print(f"[ga_mutate] bit-exact vs oracle: {np.array_equal(child, mine)}")
    # --- ES ---
    P = 6
    zs = philox.normals(1870300, philox.KIND_ES, philox.ROLE_ID[role], 5, np.arange(P), D)
    es_rows = np.zeros((P, D), dtype=np.float32)
    es_noise = np.zeros((P, len(pidx)), dtype=np.float32)
    orig_np = np.random.normal
    for i in range(P):
        agent = _load_fc(ref, env, args, role, parent)
        feeder = _NoiseFeeder(zs[i, pidx])
        np.random.normal = feeder.np_normal
        try:
            noise = agent.mutate_ES(args, role, 0, [], [], [])
        finally:
            np.random.normal = orig_np
        es_noise[i] = noise.astype(np.float32)
        es_rows[i] = layout.pack_fc_state_dict(agent.model.state_dict(), 10)
    mine_rows, mine_noise = ga_es.es_perturb(parent, args.mutation_power_agent_0, zs, pidx)
    print(f"[es_perturb] rows bit-exact: {np.array_equal(es_rows, mine_rows)}; "
          f"noise bit-exact: {np.array_equal(es_noise, mine_noise)}")
    rewards = np.linspace(-3.0, 2.0, P)
    es_mod = ref.evolutionary_strategy
    args.fitness_sharing = True
    upd_fs, div = es_mod.compute_weight_update(list(es_noise), list(rewards), args, role,
                                               individual_weights=parent[pidx],
                                               population_weights=list(es_rows[:, pidx]))
    args.fitness_sharing = False
    upd, _ = es_mod.compute_weight_update(list(es_noise), list(rewards), args, role)
    mdiv, _ = ga_es.diversity_penalty(parent[pidx], es_rows[:, pidx])
    mupd = ga_es.es_update(es_noise, rewards, args.learning_rate, args.mutation_power_agent_0)
    print(f"[es_update] max rel diff = {np.abs(mupd - upd).max() / np.abs(upd).max():.3e}; "
          f"diversity ref={div} oracle={mdiv}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "mutation.npz"),
                        source="agent.py:25-70, evolutionary_strategy.py:120-148, "
                               "utils/game_logic_functions.py:12-37",
                        parent_seed=401, ln_jitter=0.02, seed=1870300,
                        ga_gen=3, ga_member=7, ga_sigma=0.05,
                        ga_child_crc=np.array([np.bitwise_xor.reduce(child.view(np.uint32))]),
                        ga_child_head=child[:64], ga_child_tail=child[-64:],
                        es_gen=5, es_P=P, es_sigma=args.mutation_power_agent_0,
                        es_rows_crc=np.array([np.bitwise_xor.reduce(es_rows.view(np.uint32).reshape(-1))]),
                        es_rewards=rewards, lr=args.learning_rate,
                        es_update=upd, es_update_fs=upd_fs, es_diversity=div)


def golden_selection():
    """np.argsort(f)[::-1][:E] (genetic_algorithm.py:223-234) on tie-free data,
    plus the tie convention this build defines."""
    rng = np.random.Generator(np.random.PCG64(501))
    f = rng.standard_normal(1000)
    ref_idx = np.argsort(list(f))[::-1][:7]
    assert np.array_equal(ref_idx, ga_es.select_topk(f, 7))
    ft = np.round(f, 1)                       # many ties
    np.savez_compressed(os.path.join(GOLDEN_DIR, "selection.npz"),
                        source="genetic_algorithm.py:223-234", fitness=f, top7=ref_idx,
                        fitness_ties=ft, top7_ties=ga_es.select_topk(ft, 7))
    print("[selection] ok")


def golden_deepqn(ref):
    """DeepQN.forward (Atari/deepqn.py:39-48) on synthetic frames."""
    out = {}
    for c_in, n_act in ((4, 6), (4, 18), (6, 18)):
        rows = weights.make_dqn_rows(2, c_in, n_act, 600 + c_in + n_act, bn_jitter=0.05)
        rng = np.random.Generator(np.random.PCG64(77))
        frames = rng.integers(0, 256, (2, 2, c_in, 84, 84), dtype=np.uint8)
        logits = np.zeros((2, 2, n_act), dtype=np.float32)
        segs, _ = layout.dqn_segments(c_in, n_act)
        for m in range(2):
            net = ref.Atari_deepqn.DeepQN(c_in, n_act, "float32")
            sd = net.state_dict()
            for name, off, shape, _ in segs:
                sd[name] = torch.from_numpy(np.array(rows[m, off:off + int(np.prod(shape))]).reshape(shape))
            net.load_state_dict(sd)
            for f in range(2):
                with torch.no_grad():
                    x = torch.from_numpy(frames[m, f].astype(np.float32))[None]
                    logits[m, f] = net.forward(x)[0].numpy()
        mine, _ = odqn.dqn_forward_batch(rows, frames, c_in, n_act)
        print(f"[deepqn C={c_in} A={n_act}] max |oracle-ref| = {np.abs(mine - logits).max():.3e} "
              f"(|logit| ~ {np.abs(logits).mean():.3f})")
        out[f"c{c_in}a{n_act}.logits"] = logits
    np.savez_compressed(os.path.join(GOLDEN_DIR, "deepqn.npz"),
                        source="Atari/deepqn.py:39-48 DeepQN.forward (train-mode BN, batch 1)",
                        frame_seed=77, bn_jitter=0.05, **out)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = stubs.import_reference()
    torch.set_num_threads(1)
    args = stubs.RefArgs()
    golden_fc_logits(ref, args)
    golden_episodes(ref, args)
    golden_step_limit(ref)
    golden_mutation(ref, stubs.RefArgs(mutation_power_agent_0=0.05))
    golden_selection()
    golden_deepqn(ref)
    return 0


if __name__ == "__main__":
    sys.exit(main())
