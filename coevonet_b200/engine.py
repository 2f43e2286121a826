"""Device-resident Co-GA / Co-ES engines behind the reference's train loops.

One process per GPU.  The population of each role is a ``float32[n_local,
pitch]`` tensor holding the contiguous row block ``[row0, row0 + n_local)`` of
this rank; Hall-of-Fame rows, ES base vectors and all
scalar state are replicated.  Collectives (NCCL on GPUs, gloo in the CPU
tests) appear only where the path has an exchange step:

* all-gather of the per-member fitness (and fitness-sharing distances),
* an all-reduce that assembles the elite rows on every rank (each elite row is
  contributed by its owner, zeros elsewhere) -- this is the HoF broadcast,
* the all-reduce of the ES update ``delta``.

Everything numeric goes through ``self.k`` -- by default ``coevonet_b200.ops``
(the CUDA kernels).  Tests may inject a checker backend with the same call
signatures to exercise the sharding logic without a GPU; the product never does.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import layout
from .utils import mpe_spec

ROLES = layout.ROLES                    # ("agent_0", "agent_1", "adversary_0")
N_EVAL_GAMES = 10                       # genetic_algorithm.py:18, evolutionary_strategy.py:28


# ---------------------------------------------------------------------------
# sharding + collectives
# ---------------------------------------------------------------------------
class Shard:
    """Contiguous block partition of P rows over `world` ranks."""

    def __init__(self, P, rank=0, world=1):
        self.P, self.rank, self.world = int(P), int(rank), int(world)
        base, rem = divmod(self.P, self.world)
        self.counts = [base + (1 if r < rem else 0) for r in range(self.world)]
        self.offsets = [sum(self.counts[:r]) for r in range(self.world)]
        self.row0 = self.offsets[self.rank]
        self.n_local = self.counts[self.rank]

    def owner(self, row):
        for r in range(self.world):
            if row < self.offsets[r] + self.counts[r]:
                return r
        raise IndexError(row)


class Comm:
    """torch.distributed when initialised with world_size > 1, otherwise a no-op."""

    def __init__(self, group=None):
        self.enabled = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.group = group
        self.rank = dist.get_rank(group) if self.enabled else 0
        self.world = dist.get_world_size(group) if self.enabled else 1

    def all_gather_rows(self, local, shard):
        """local [n_local, ...] -> [P, ...] in global row order (uneven shards are padded)."""
        if not self.enabled:
            return local
        nmax = max(shard.counts)
        pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
        bufs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(bufs, pad, group=self.group)
        return torch.cat([bufs[r][:shard.counts[r]] for r in range(self.world)], dim=0)

    def all_reduce_sum(self, t):
        if self.enabled:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def default_kernels():
    from . import ops
    return ops


def _limit_cycles(k, limit):
    return k.cycles_for_limit(limit)


# ---------------------------------------------------------------------------
# shared evaluation plumbing
# ---------------------------------------------------------------------------
class _EngineBase:
    def __init__(self, args, device, env=None, kernels=None, comm=None):
        self.args = args
        self.device = torch.device(device)
        self.k = kernels if kernels is not None else default_kernels()
        self.comm = comm if comm is not None else Comm()
        self.env = env if env is not None else mpe_spec.DeviceMPEEnv()
        self.seed = int(getattr(args, "seed", mpe_spec.ENV_SEED))
        self.compat = bool(getattr(args, "reference_compat", True))
        self.E = int(getattr(args, "envs_per_member", 1))
        self.pos_first = bool(getattr(args, "integrate_pos_first", True))
        # "reference": the host PCG64 stream env.reset() consumes upstream (exact episode-for-
        # episode initial states of a reference run); "device": Philox records generated on
        # the GPU (no host RNG work or upload; needed at large P)
        self.init_mode = getattr(args, "init_states", "reference")
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.gen = 0
        #: per-generation records (small runs / tests): filled when args.record_history is set
        self.history = [] if getattr(args, "record_history", False) else None
        #: bench hook: when a list, (start, stop) CUDA events are recorded around every
        #: population K1 launch on the launching stream
        self.k1_events = None
        #: evaluate the three roles of a generation on three CUDA streams (they are independent: every
        #: member plays the OTHER roles' base policies of the previous generation), so one role's kernels
        #: fill the partial waves and launch gaps of another's
        self.overlap_roles = bool(getattr(args, "overlap_roles", True)) and self.device.type == "cuda"
        self._role_streams = None
        #: ES update from the materialised members (K6 as an HBM-bound read) instead of regenerating the
        #: noise from the Philox key (K6 as ALU work); both are within fp32 rounding of each other
        self.update_from_members = bool(getattr(args, "update_from_members", True))

    # initial states of `n_rows` x K x E episodes for the rows [row0, row0+n_local) of a
    # P-row evaluation; in reference mode every rank draws the whole block to keep the
    # host stream in step, then slices its rows.
    def _initial_states(self, P, K, shard, stream_tag):
        E = self.E
        if self.init_mode == "reference":
            block = self.env.draw_initial_states(P * K * E).reshape(P, K, E, mpe_spec.INIT_STATE_DIM)
            mine = block[shard.row0:shard.row0 + shard.n_local]
            return torch.from_numpy(np.ascontiguousarray(mine)).to(self.device)
        rec0 = shard.row0 * K * E
        out = self.k.init_states(self.seed, stream_tag, shard.n_local * K * E, self.device, rec0=rec0)
        return out.reshape(shard.n_local, K, E, mpe_spec.INIT_STATE_DIM)

    def _check_status(self):
        self.k.raise_on_status(self.status)

    def _role_slot(self, out, role, limit):
        s0, s1, sadv = self.k.reward_slots(out, agent_step_limit=limit, reference_compat=self.compat)
        return {"agent_0": s0, "agent_1": s1, "adversary_0": sadv}[role]

    def evaluate_triple(self, row_a0, row_a1, row_adv):
        """``evaluate_current_weights`` (genetic_algorithm.py:12-29): mean reward
        triple of 10 eval games between three single rows.  Replicated on every rank."""
        limit = self.args.max_evaluation_steps
        if self.init_mode == "reference":
            init = torch.from_numpy(self.env.draw_initial_states(N_EVAL_GAMES)).to(self.device)
        else:
            init = self.k.init_states(self.seed, 0x40000000 + self.gen, N_EVAL_GAMES, self.device)
        init = init.reshape(1, 1, N_EVAL_GAMES, mpe_spec.INIT_STATE_DIM)
        out = self.k.mpe_rollout("agent_0", row_a0.reshape(1, -1), row_adv.reshape(1, -1),
                                 row_a1.reshape(1, -1), init, n_cycles=_limit_cycles(self.k, limit),
                                 pos_first=self.pos_first, status=self.status)
        s0, s1, sadv = self.k.reward_slots(out, agent_step_limit=limit, reference_compat=self.compat)
        return float(s0.mean()), float(s1.mean()), float(sadv.mean())


# ---------------------------------------------------------------------------
# Co-GA  (genetic_algorithm.py:51-374)
# ---------------------------------------------------------------------------
class GAEngine(_EngineBase):
    """State: per role the sharded population, the replicated HoF ring
    (oldest first, like the reference's list) and the frozen founder row the
    reference measures diversity against (Appendix C #3)."""

    def __init__(self, args, device, pop_rows, hof_rows, founder_rows, env=None, kernels=None, comm=None):
        super().__init__(args, device, env, kernels, comm)
        self.P = int(args.population)
        self.shard = Shard(self.P, self.comm.rank, self.comm.world)
        self.pop = {r: pop_rows[r].to(self.device).contiguous() for r in ROLES}       # [n_local, pitch]
        self.hof = {r: hof_rows[r].to(self.device).contiguous() for r in ROLES}       # [hof, pitch]
        self.founder = {r: founder_rows[r].to(self.device).contiguous() for r in ROLES}
        self.elites = {r: None for r in ROLES}
        self.elite_ids = {r: None for r in ROLES}
        self.fitness = {r: None for r in ROLES}
        self.diversity = {r: None for r in ROLES}
        for r in ROLES:
            assert self.pop[r].shape[0] == self.shard.n_local, "population shard has the wrong row count"

    def sigma(self, role):
        a = self.args
        return {"agent_0": a.mutation_power_agent_0, "agent_1": a.mutation_power_agent_1,
                "adversary_0": a.mutation_power_adversary}[role]

    def _opponents(self, role):
        """Rows for the two other seats (ascending seat order), newest HoF entry
        first (``hof[len-1-k]``, genetic_algorithm.py:138-139,170-171,203-204)."""
        newest_first = {r: torch.flip(self.hof[r], dims=[0]).contiguous() for r in ROLES}
        if role == "agent_0":       # seats: adversary_0, agent_1
            return newest_first["adversary_0"], newest_first["agent_1"]
        if role == "agent_1":       # seats: adversary_0, agent_0
            return newest_first["adversary_0"], newest_first["agent_0"]
        # adversary: seats agent_0, agent_1; the reference seats hof_agent_0 in BOTH chairs
        # (genetic_algorithm.py:204, Appendix C #5)
        second = newest_first["agent_0"] if self.compat else newest_first["agent_1"]
        return newest_first["agent_0"], second

    def evaluate_role(self, role):
        """Evaluation loop of one role (genetic_algorithm.py:125-217) -> global fitness fp64[P]."""
        a = self.args
        in_dim = layout.OBS_DIM[role]
        K = int(a.hof_size)
        limit = a.max_timesteps_per_episode
        opp_a, opp_b = self._opponents(role)
        init = self._initial_states(self.P, K, self.shard, ROLES.index(role) + 4 * self.gen)
        if self.compat and not getattr(a, "play_discarded_hof_games", False):
            # The reference plays hof_size games per member but OVERWRITES the reward each time
            # (`=`, genetic_algorithm.py:140,172,205; Appendix C #2): only the game against the
            # oldest HoF entry (k = K-1) reaches the fitness.  The discarded games are not
            # simulated here; their initial-state records are still drawn (above), so the game
            # that counts starts from the same state as in a reference run.
            opp_a, opp_b = opp_a[K - 1:K].contiguous(), opp_b[K - 1:K].contiguous()
            init = init[:, K - 1:K].contiguous()
        out = self.k.mpe_rollout(role, self.pop[role], opp_a, opp_b, init, n_cycles=_limit_cycles(self.k, limit),
                                 pos_first=self.pos_first, status=self.status)
        slot = self._role_slot(out, role, limit)                      # [n_local, K or 1, E]
        if self.compat:
            # only the LAST HoF game counts (reward overwritten, `=`), then / hof_size
            # (genetic_algorithm.py:140-144, Appendix C #2)
            reward = slot[:, -1, :].mean(dim=1) / K
        else:
            reward = slot.mean(dim=(1, 2))
        # fitness sharing against the frozen founder (Appendix C #3); applied regardless of
        # --fitness_sharing in GA (Appendix C #4)
        dist_local = self.k.diversity_dist(self.pop[role], self.founder[role], in_dim)
        dist_all = self.comm.all_gather_rows(dist_local, self.shard)
        div = self.k.diversity_from_dist(dist_all)
        self.diversity[role] = float(div)
        if self.compat or a.fitness_sharing:
            reward = reward / (1.0 + div.to(reward.dtype))
        fit = self.comm.all_gather_rows(reward.contiguous(), self.shard)
        self.fitness[role] = fit
        return fit

    def select_and_repopulate(self, role):
        """Truncation selection, HoF FIFO update, elite cloning + mutation
        (genetic_algorithm.py:223-290)."""
        a = self.args
        in_dim = layout.OBS_DIM[role]
        E = int(a.elites_number)
        ids = self.k.select_topk(self.fitness[role], E)                       # replicated, bit-exact
        self.elite_ids[role] = ids
        ids_host = ids.cpu().tolist()
        pitch = self.pop[role].shape[1]
        elites = torch.zeros((E, pitch), dtype=torch.float32, device=self.device)
        mine = [(j, g - self.shard.row0) for j, g in enumerate(ids_host)
                if self.shard.row0 <= g < self.shard.row0 + self.shard.n_local]
        if mine:
            local_idx = torch.tensor([m[1] for m in mine], dtype=torch.int64, device=self.device)
            rows = self.k.gather_rows(self.pop[role], local_idx)
            elites[torch.tensor([m[0] for m in mine], device=self.device)] = rows
        self.comm.all_reduce_sum(elites)                                      # HoF / elite broadcast
        self.elites[role] = elites
        # HoF: append the best, drop the oldest (genetic_algorithm.py:270-275)
        self.hof[role] = torch.cat([self.hof[role][1:], elites[0:1]], dim=0).contiguous()
        # next population: row 0 = best unmutated, rows c>=1 = elites[(c-1)%E] + sigma*N(0,1)
        self.k.ga_repopulate(elites, layout.fc_dim(in_dim), self.sigma(role), self.seed, role, self.gen,
                             self.shard.row0, self.shard.n_local, out=self.pop[role])

    def step(self):
        """One generation up to (not including) the host-side sigma adaptation."""
        for role in ROLES:
            self.evaluate_role(role)
        self._check_status()
        for role in ROLES:
            self.select_and_repopulate(role)
        best = {r: self.elites[r][0] for r in ROLES}
        ev = self.evaluate_triple(best["agent_0"], best["agent_1"], best["adversary_0"])
        self._check_status()
        if self.history is not None:
            self.history.append(dict(
                fitness={r: self.fitness[r].cpu().numpy().copy() for r in ROLES},
                elite_ids={r: self.elite_ids[r].cpu().numpy().copy() for r in ROLES},
                diversity=dict(self.diversity), evals=ev,
                sigma={r: self.sigma(r) for r in ROLES}))
        self.gen += 1
        return ev


# ---------------------------------------------------------------------------
# Co-ES  (evolutionary_strategy.py:151-393)
# ---------------------------------------------------------------------------
class ESEngine(_EngineBase):
    """State: one replicated base row per role.  Members of a generation are
    ``theta + sigma * N(0,1)`` regenerated from the Philox key both when they are
    materialised for the rollout (K5) and when the update is formed (K6)."""

    def __init__(self, args, device, theta_rows, env=None, kernels=None, comm=None):
        super().__init__(args, device, env, kernels, comm)
        self.P = int(args.population)
        self.shard = Shard(self.P, self.comm.rank, self.comm.world)
        self.theta = {r: theta_rows[r].to(self.device).contiguous().reshape(-1) for r in ROLES}
        self.members = {r: torch.empty((self.shard.n_local, layout.fc_pitch(layout.OBS_DIM[r])),
                                       dtype=torch.float32, device=self.device) for r in ROLES}
        self.fitness = {r: None for r in ROLES}
        self.rewards = {r: None for r in ROLES}
        self.diversity = {r: None for r in ROLES}
        self.last_delta = {r: None for r in ROLES}

    def sigma(self, role):
        a = self.args
        return {"agent_0": a.mutation_power_agent_0, "agent_1": a.mutation_power_agent_1,
                "adversary_0": a.mutation_power_adversary}[role]

    def _base_opponents(self, role):
        t = self.theta
        if role == "agent_0":
            return t["adversary_0"].reshape(1, -1), t["agent_1"].reshape(1, -1)
        if role == "agent_1":
            return t["adversary_0"].reshape(1, -1), t["agent_0"].reshape(1, -1)
        return t["agent_0"].reshape(1, -1), t["agent_1"].reshape(1, -1)

    def _reference_initial_states(self):
        """The reference interleaves the three roles per member
        (evolutionary_strategy.py:236-251): game 3*i + role_index."""
        E = self.E
        block = self.env.draw_initial_states(self.P * 3 * E).reshape(self.P, 3, E, mpe_spec.INIT_STATE_DIM)
        sl = slice(self.shard.row0, self.shard.row0 + self.shard.n_local)
        return {r: torch.from_numpy(np.ascontiguousarray(block[sl, i][:, None])).to(self.device)
                for i, r in enumerate(ROLES)}

    def evaluate(self, init_by_role=None):
        """Perturb + one rollout per member and role against the other roles'
        BASE policies (mutate_weights, evolutionary_strategy.py:63-116).
        ``init_by_role`` (optional) supplies the initial states [n_local,1,E,11] per role."""
        limit = self.args.max_timesteps_per_episode
        ref_init = init_by_role
        if ref_init is None and self.init_mode == "reference":
            ref_init = self._reference_initial_states()
        def one_role(role):
            in_dim = layout.OBS_DIM[role]
            self.k.es_perturb(self.theta[role], in_dim, self.sigma(role), self.seed, role, self.gen,
                              self.shard.row0, self.shard.n_local, out=self.members[role])
            if ref_init is not None:
                init = ref_init[role]
            else:
                init = self._initial_states(self.P, 1, self.shard, ROLES.index(role) + 4 * self.gen)
            opp_a, opp_b = self._base_opponents(role)
            if self.k1_events is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            out = self.k.mpe_rollout(role, self.members[role], opp_a, opp_b, init,
                                     n_cycles=_limit_cycles(self.k, limit), pos_first=self.pos_first,
                                     status=self.status)
            if self.k1_events is not None:
                e1.record()
                self.k1_events.append((e0, e1))
            slot = self._role_slot(out, role, limit)                  # [n_local, 1, E]
            self.rewards[role] = slot.mean(dim=(1, 2)).contiguous()

        if self.overlap_roles and self.k1_events is None:
            if self._role_streams is None:
                self._role_streams = [torch.cuda.Stream(device=self.device) for _ in ROLES]
            main = torch.cuda.current_stream(self.device)
            for role, st in zip(ROLES, self._role_streams):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    one_role(role)
            for st in self._role_streams:
                main.wait_stream(st)
        else:
            for role in ROLES:
                one_role(role)

    def update(self):
        """compute_weight_update + apply (evolutionary_strategy.py:120-148,255-265)."""
        a = self.args
        for role in ROLES:
            in_dim = layout.OBS_DIM[role]
            # the reference casts rewards to fp32 before the update (np.array(rewards, dtype=np_dtype))
            fit_local = self.rewards[role].to(torch.float32).to(torch.float64)
            if a.fitness_sharing:
                dist_local = self.k.diversity_dist(self.members[role], self.theta[role], in_dim)
                dist_all = self.comm.all_gather_rows(dist_local, self.shard)
                div = self.k.diversity_from_dist(dist_all)
                self.diversity[role] = float(div)
                fit_local = (fit_local.to(torch.float32) / (1 + div)).to(torch.float64)
            self.fitness[role] = self.comm.all_gather_rows(fit_local.contiguous(), self.shard)
            if self.update_from_members and hasattr(self.k, "es_update_members"):
                # sigma*z_i read back from the materialised members (HBM bound) instead of regenerated
                delta = self.k.es_update_members(fit_local.contiguous(), self.members[role], self.theta[role],
                                                 in_dim, self.sigma(role), a.learning_rate, self.P)
            else:
                delta = self.k.es_update(fit_local.contiguous(), in_dim, self.sigma(role), a.learning_rate, self.P,
                                         self.seed, role, self.gen, self.shard.row0)
            self.comm.all_reduce_sum(delta)
            self.k.axpy(1.0, delta, self.theta[role])
            self.last_delta[role] = delta

    def step(self, init_by_role=None):
        self.evaluate(init_by_role)
        self._check_status()
        self.update()
        ev = self.evaluate_triple(self.theta["agent_0"], self.theta["agent_1"], self.theta["adversary_0"])
        self._check_status()
        if self.history is not None:
            rewards_all = {r: self.comm.all_gather_rows(self.rewards[r], self.shard).cpu().numpy().copy()
                           for r in ROLES}
            self.history.append(dict(
                rewards=rewards_all, diversity=dict(self.diversity), evals=ev,
                delta={r: self.last_delta[r].cpu().numpy().copy() for r in ROLES},
                sigma={r: self.sigma(r) for r in ROLES}))
        self.gen += 1
        return ev


def adapt_sigma(args, hist0, hist1, histadv, gen):
    """Dynamic mutation power (genetic_algorithm.py:323-345 ==
    evolutionary_strategy.py:292-316), including agent_0 growing from
    sigma_agent_1 * 1.2 (Appendix C #6).  Host scalars, mutates ``args``."""
    def worse(h):
        return gen > 10 and np.mean(h[-10:]) < np.mean(h[-20:-10])
    if worse(hist0):
        args.mutation_power_agent_0 = min(args.mutation_power_agent_1 * 1.2, args.max_mutation_power)
    else:
        args.mutation_power_agent_0 = max(args.mutation_power_agent_0 * 0.95, args.min_mutation_power)
    if worse(hist1):
        args.mutation_power_agent_1 = min(args.mutation_power_agent_1 * 1.2, args.max_mutation_power)
    else:
        args.mutation_power_agent_1 = max(args.mutation_power_agent_1 * 0.95, args.min_mutation_power)
    if worse(histadv):
        args.mutation_power_adversary = min(args.mutation_power_adversary * 1.2, args.max_mutation_power)
    else:
        args.mutation_power_adversary = max(args.mutation_power_adversary * 0.95, args.min_mutation_power)
