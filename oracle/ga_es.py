"""CPU restatement of the GA / ES population arithmetic (K3-K7's checker).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function cites the
reference lines it follows; pinned against the reference's own functions by
``oracle/make_golden.py``.
"""
from __future__ import annotations

import numpy as np


def select_topk(fitness, k):
    """Indices of the ``k`` largest fitness values, descending.

    Follows genetic_algorithm.py:223-234 (``np.argsort(f)[::-1][:E]``).  The
    reference's sort is unstable, so ties are implementation-defined there;
    this build DEFINES ties -> lower index first which
    is ``argsort(-f, kind='stable')``.  NaN cannot occur (forward raises)."""
    f = np.asarray(fitness, dtype=np.float64)
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
