"""CPU restatement of the GA / ES population arithmetic (K3-K7's checker).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function cites the
reference lines it follows; pinned against the reference's own functions by
``oracle/make_golden.py``.
"""
from __future__ import annotations

import numpy as np


def select_topk(fitness, k, order=0):
    """Indices of the ``k`` largest fitness values, descending.

    Follows genetic_algorithm.py:223-234 (``np.argsort(f)[::-1][:E]``).  The
    reference's default sort is unstable for large arrays, so ties are
    implementation-defined there.  ``order=0`` DEFINES ties -> lower index first
     = ``argsort(-f, kind='stable')``, NaN last.
    ``order=1`` is the reference's expression with a STABLE ascending sort --
    what NumPy's insertion sort does for n <= 16, the sizes of the reference's own
    runs: ties -> higher index first, NaN first (Appendix C #18)."""
    f = np.asarray(fitness, dtype=np.float64)
    if order == 1:
        return np.argsort(f, kind="stable")[::-1][:k].astype(np.int64)
    return np.argsort(-f, kind="stable")[:k].astype(np.int64)


def ga_repopulate(pop, elite_idx, sigma, z):
    """Next GA population (genetic_algorithm.py:255-290 + mutate_elites :32-48
    + MPEAgent.clone MPE/mpe_agent.py:24-28 + Agent.mutate agent.py:25-29).

    row 0          = the best member, unmutated (``population = [best]``)
    row c (c >= 1) = elites[(c-1) % E] + sigma * z[c]   (EVERY parameter,
                     LayerNorm gamma/beta included)
    ``z`` is fp32[P, D] standard normals (row 0 unused); arithmetic is two
    separately rounded fp32 ops (mul, add) like ``param.data += noise``.
    """
    pop = np.asarray(pop, dtype=np.float32)
    elite_idx = np.asarray(elite_idx, dtype=np.int64)
    P, D = pop.shape
    E = len(elite_idx)
    out = np.empty_like(pop)
    out[0] = pop[elite_idx[0]]
    sig = np.float32(sigma)
    for c in range(1, P):
        parent = pop[elite_idx[(c - 1) % E]]
        out[c] = parent + (sig * z[c, :D].astype(np.float32)).astype(np.float32)
    return out


def ga_repopulate_crossover(elites, sigma, z, crossover_rate, seed, role_id, gen, members):
    """Rows ``members`` (global ids) of the next population WITH the crossover extension (the
    reference has none: README.md:47 vs genetic_algorithm.py:32-48).

    child c >= 1: parent A = elites[(c-1) % E]; Philox(seed, XOVER, role, gen, member=c) block
    0xFFFFFFFF: word0 < rate * 2^32 -> recombine with mate B = elites[((c-1) % E + 1 + word1 % (E-1)) % E];
    block j4: bit 0 of word i set -> parameter 4*j4+i from A, else from B.  Then + sigma * z[c].
    ``z`` fp32 [len(members), D]."""
    from . import philox
    elites = np.asarray(elites, dtype=np.float32)
    E, D = elites.shape
    out = np.empty((len(members), D), dtype=np.float32)
    sig = np.float32(sigma)
    thr = float(np.float32(crossover_rate)) * 4294967296.0
    n4 = (D + 3) // 4
    for r, c in enumerate(members):
        c = int(c)
        if c == 0:
            out[r] = elites[0]
            continue
        a = (c - 1) % E
        row = elites[a].copy()
        if crossover_rate > 0 and E >= 2:
            k0, k1 = philox.split_seed(seed)
            tag = np.uint32((role_id & 0xFF) | (philox.KIND_XOVER << 8))
            y = philox.philox4x32_10(np.uint32(0xFFFFFFFF), np.uint32(c), np.uint32(gen), tag, k0, k1)
            if float(y[0]) < thr:
                b = (a + 1 + int(y[1]) % (E - 1)) % E
                w = philox.words(seed, philox.KIND_XOVER, role_id, gen, [c], n4).reshape(-1)[:D]
                take_a = (w & np.uint32(1)).astype(bool)
                row = np.where(take_a, elites[a], elites[b]).astype(np.float32)
        out[r] = row + (sig * z[r, :D].astype(np.float32)).astype(np.float32)
    return out


def hof_update(hof, best_row):
    """FIFO Hall of Fame (genetic_algorithm.py:270-275): append newest, drop
    oldest; opponents are read newest-first ``hof[len-1-k]`` (:138-139)."""
    hof = np.asarray(hof)
    return np.concatenate([hof[1:], np.asarray(best_row)[None]], axis=0)


def diversity_penalty(individual, population):
    """utils/game_logic_functions.py:12-37 in the reference's dtypes (fp32
    inputs => fp32 norms; sigma = mean distance)."""
    population = np.asarray(population)
    individual = np.asarray(individual)
    d = np.array([np.linalg.norm(p - individual) for p in population])
    sigma = np.mean(d)
    return float(np.sum(np.maximum(0, 1 - d / sigma))), d


def es_perturb(theta, sigma, z, pert_idx):
    """Perturbed member rows (agent.py:31-70): theta + sigma*z on the
    perturbable entries only (norm layers untouched).  Returns (rows, noise)
    where ``noise = sigma*z`` is what ``mutate_ES`` returns (already scaled,
    Appendix C #13)."""
    theta = np.asarray(theta, dtype=np.float32)
    z = np.asarray(z, dtype=np.float32)
    P = z.shape[0]
    rows = np.repeat(theta[None], P, axis=0)
    noise = (np.float32(sigma) * z[:, pert_idx]).astype(np.float32)
    rows[:, pert_idx] = theta[pert_idx] + noise
    return rows, noise


def es_update(noises, rewards, lr, sigma, diversity=None):
    """compute_weight_update (evolutionary_strategy.py:120-148), fp32:
    delta = lr / (n * sigma) * noises.T @ fitness, fitness = rewards/(1+div)."""
    noises = np.asarray(noises, dtype=np.float32)
    rewards = np.asarray(rewards, dtype=np.float32)
    fitness = rewards / (1 + diversity) if diversity is not None else rewards
    upd = (lr / (len(noises) * sigma)) * np.dot(noises.T, fitness)
    return upd.astype(np.float32)


def adaptive_sigma(sig0, sig1, sigadv, hist0, hist1, histadv, gen, smin, smax):
    """genetic_algorithm.py:323-345 == evolutionary_strategy.py:292-316,
    including quirk #6 (agent_0 grows from sigma_agent_1 * 1.2)."""
    def worse(h):
        return gen > 10 and np.mean(h[-10:]) < np.mean(h[-20:-10])
    if worse(hist0):
        sig0 = min(sig1 * 1.2, smax)
    else:
        sig0 = max(sig0 * 0.95, smin)
    if worse(hist1):
        sig1 = min(sig1 * 1.2, smax)
    else:
        sig1 = max(sig1 * 0.95, smin)
    if worse(histadv):
        sigadv = min(sigadv * 1.2, smax)
    else:
        sigadv = max(sigadv * 0.95, smin)
    return sig0, sig1, sigadv


def reward_slots(out, agent_step_limit=None, reference_compat=True):
    """play_MPE's three return slots from the physical sums (sum_good, last_good, sum_adv)
    (utils/game_logic_functions.py:179-190)."""
    sg, lg, sa = out[..., 0], out[..., 1], out[..., 2]
    if not reference_compat:
        return sg, sg, sa
    L = 75 if agent_step_limit is None else int(min(75, max(0, int(agent_step_limit))))
    nc = L // 3
    if nc == 0:
        z = np.zeros_like(sg)
        return z, sa, z
    n_adv = -(-L // 3)
    n_a0 = -(-(L - 1) // 3)
    slot_adv = sg if n_adv - 1 >= nc else sg - lg
    slot_a0 = sg if n_a0 - 1 >= nc else sg - lg
    return slot_a0, sa, slot_adv


# slots of the device generation state (CEV_GS_* in include/coevonet_b200.h)
GS_GEN, GS_SIGMA, GS_BEST, GS_STALE, GS_STOP, GS_STOP_GEN, GS_LAST_EVAL, GS_HIST = 0, 1, 4, 7, 10, 11, 12, 16


def generation_state(sigmas, hist_capacity):
    gs = np.zeros(GS_HIST + 3 * hist_capacity + 3 * (hist_capacity + 1))
    gs[GS_SIGMA:GS_SIGMA + 3] = sigmas
    gs[GS_BEST:GS_BEST + 3] = -np.inf
    gs[GS_HIST + 3 * hist_capacity:GS_HIST + 3 * hist_capacity + 3] = sigmas
    return gs


def generation_end(eval_out, gs, hist_capacity, agent_step_limit=None, reference_compat=True, adaptive=False,
                   sigma_max=0.5, sigma_min=0.001, early_stopping=False, min_delta=0.1, patience=300):
    """Tail of a training-loop iteration, restated with the reference's own expressions:
    evaluate_current_weights' running sums / 10 (genetic_algorithm.py:12-29), the reward lists,
    the adaptive mutation power (genetic_algorithm.py:323-345 == evolutionary_strategy.py:292-316,
    np.mean on Python-float lists) and the early-stopping counters (evolutionary_strategy.py:318-354).
    ``gs`` is updated in place (same slots as the device array)."""
    out = np.asarray(eval_out, dtype=np.float64).reshape(-1, 4)
    gen = int(gs[GS_GEN])
    tot = [0.0, 0.0, 0.0]
    for g in range(out.shape[0]):
        s0, s1, sadv = reward_slots(out[g], agent_step_limit, reference_compat)
        tot[0] += float(s0)
        tot[1] += float(s1)
        tot[2] += float(sadv)
    ev = [t / out.shape[0] for t in tot]
    hist = gs[GS_HIST:GS_HIST + 3 * hist_capacity].reshape(hist_capacity, 3)
    shist = gs[GS_HIST + 3 * hist_capacity:].reshape(hist_capacity + 1, 3)
    if gen < hist_capacity:
        hist[gen] = ev
    gs[GS_LAST_EVAL:GS_LAST_EVAL + 3] = ev
    if adaptive:
        lists = [list(hist[:gen + 1, r]) for r in range(3)]
        sig = adaptive_sigma(gs[GS_SIGMA], gs[GS_SIGMA + 1], gs[GS_SIGMA + 2], lists[0], lists[1], lists[2], gen,
                             sigma_min, sigma_max)
        gs[GS_SIGMA:GS_SIGMA + 3] = sig
    if gen + 1 <= hist_capacity:
        shist[gen + 1] = gs[GS_SIGMA:GS_SIGMA + 3]
    if early_stopping and gs[GS_STOP] == 0:
        for r in range(3):
            if ev[r] > gs[GS_BEST + r] + min_delta:
                gs[GS_BEST + r] = ev[r]
                gs[GS_STALE + r] = 0
            else:
                gs[GS_STALE + r] += 1
        for r in range(3):
            if gs[GS_STALE + r] >= patience:
                gs[GS_STOP] = 1 + r
                gs[GS_STOP_GEN] = gen
                break
    gs[GS_GEN] = gen + 1
    return gs


def weight_stats(rows, pert_idx):
    """MPEAgent.log_weight_statistics (MPE/mpe_agent.py:30-50): mean, min, max, std (ddof 0)
    of every row's perturbable weights."""
    w = np.asarray(rows, dtype=np.float32)[:, pert_idx]
    return np.stack([w.mean(axis=1), w.min(axis=1), w.max(axis=1), w.std(axis=1)], axis=1)
