"""NumPy restatement of PettingZoo ``simple_adversary_v3`` (AEC API).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: the
upstream package (``pettingzoo``, unpinned, ``/root/reference/requirements.txt:5``;
1.24.x assumed) is absent from the reference tree and from this image, so this
file restates its published algorithm (``pettingzoo/mpe/simple_adversary/
simple_adversary.py``, ``pettingzoo/mpe/_mpe_utils/{core,simple_env}.py``) as
summarised in SURVEY.md Appendix A.  It exposes exactly the AEC surface the
reference calls:

* ``env(render_mode=None)``            utils/game_logic_functions.py:46
* ``reset(seed=...)`` / ``reset()``    utils/game_logic_functions.py:54,217
* ``agent_iter()``                     utils/game_logic_functions.py:130
* ``observe(agent)``                   utils/game_logic_functions.py:138
* ``step(action)``                     utils/game_logic_functions.py:179
* ``last()``                           utils/game_logic_functions.py:181
* ``observation_space(a).shape`` / ``action_space(a).n``   MPE/mpe_agent.py:17-18

One upstream-version-dependent choice is a switch: ``INTEGRATE_POS_FIRST``
(Appendix A.4).  It is recorded with every golden vector.
"""
from __future__ import annotations

import numpy as np

# --- world constants (Appendix A.1/A.2) -------------------------------------
DT = 0.1
DAMPING = 0.25
SENSITIVITY = 5.0
MASS = 1.0
MAX_CYCLES = 25
AGENT_ORDER = ("adversary_0", "agent_0", "agent_1")
OBS_DIM = {"adversary_0": 8, "agent_0": 10, "agent_1": 10}
N_ACTIONS = 5

#: PettingZoo >= 1.24 integrates position before velocity; older releases and
#: the original OpenAI MPE integrate velocity first (Appendix A.4).
INTEGRATE_POS_FIRST = True

#: columns of the flat initial-state record used across the whole repo
#: [goal_idx, adv.xy, agent_0.xy, agent_1.xy, landmark0.xy, landmark1.xy]
INIT_STATE_DIM = 11


class _Space:
    def __init__(self, shape=None, n=None):
        self.shape = shape
        self.n = n


class _Landmark:
    """Opaque landmark object (``np_random.choice`` is called on a list of
    these, exactly like upstream ``reset_world``)."""

    def __init__(self, i):
        self.i = i
        self.p_pos = np.zeros(2)


class SimpleAdversaryEnv:
    """AEC environment with the upstream turn/reward bookkeeping (A.7)."""

    metadata = {"name": "simple_adversary_v3", "is_parallelizable": True}

    def __init__(self, render_mode=None, max_cycles=MAX_CYCLES,
                 integrate_pos_first=None):
        self.render_mode = render_mode
        self.max_cycles = max_cycles
        self.integrate_pos_first = (INTEGRATE_POS_FIRST if integrate_pos_first is None
                                    else bool(integrate_pos_first))
        self.possible_agents = list(AGENT_ORDER)
        self.agents = list(AGENT_ORDER)
        self._index = {a: i for i, a in enumerate(AGENT_ORDER)}
        self.landmarks = [_Landmark(0), _Landmark(1)]
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence()))
        self.p_pos = np.zeros((3, 2))
        self.p_vel = np.zeros((3, 2))
        self.goal = 0
        self.steps = 0
        self._sel = 0
        self.agent_selection = AGENT_ORDER[0]
        self.current_actions = [0, 0, 0]
        self.rewards = {a: 0.0 for a in AGENT_ORDER}
        self._cumulative_rewards = {a: 0.0 for a in AGENT_ORDER}
        self.terminations = {a: False for a in AGENT_ORDER}
        self.truncations = {a: False for a in AGENT_ORDER}
        self.infos = {a: {} for a in AGENT_ORDER}
        #: every reset appends the flat initial-state record here (test hook)
        self.init_state_log = []

    # -- spaces ---------------------------------------------------------------
    def observation_space(self, agent):
        return _Space(shape=(OBS_DIM[agent],))

    def action_space(self, agent):
        return _Space(n=N_ACTIONS)

    # -- reset (A.3) ----------------------------------------------------------
    def reset(self, seed=None, options=None):
        if seed is not None:
            # gymnasium.utils.seeding.np_random(seed)
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        goal = self.np_random.choice(self.landmarks)
        self.goal = goal.i
        for i in range(3):
            self.p_pos[i] = self.np_random.uniform(-1, +1, 2)
            self.p_vel[i] = 0.0
        for lm in self.landmarks:
            lm.p_pos = self.np_random.uniform(-1, +1, 2)
        self.agents = list(AGENT_ORDER)
        self.rewards = {a: 0.0 for a in AGENT_ORDER}
        self._cumulative_rewards = {a: 0.0 for a in AGENT_ORDER}
        self.terminations = {a: False for a in AGENT_ORDER}
        self.truncations = {a: False for a in AGENT_ORDER}
        self.infos = {a: {} for a in AGENT_ORDER}
        self._sel = 0
        self.agent_selection = AGENT_ORDER[0]
        self.steps = 0
        self.current_actions = [0, 0, 0]
        self.init_state_log.append(self.flat_state())

    def flat_state(self):
        """[goal, adv.xy, a0.xy, a1.xy, lm0.xy, lm1.xy] as float64[11]."""
        return np.concatenate([[float(self.goal)], self.p_pos.reshape(-1),
                               self.landmarks[0].p_pos, self.landmarks[1].p_pos])

    def load_flat_state(self, rec):
        """Inverse of :meth:`flat_state` (velocities zero, fresh episode)."""
        self.reset()
        self.init_state_log.pop()
        rec = np.asarray(rec, dtype=np.float64)
        self.goal = int(rec[0])
        self.p_pos[:] = rec[1:7].reshape(3, 2)
        self.p_vel[:] = 0.0
        self.landmarks[0].p_pos = rec[7:9].copy()
        self.landmarks[1].p_pos = rec[9:11].copy()

    # -- iteration ------------------------------------------------------------
    def agent_iter(self, max_iter=2 ** 63):
        n = 0
        while self.agents and n < max_iter:
            yield self.agent_selection
            n += 1

    # -- observation (A.6) ----------------------------------------------------
    def observe(self, agent):
        i = self._index[agent]
        me = self.p_pos[i]
        entity_pos = [lm.p_pos - me for lm in self.landmarks]
        other_pos = [self.p_pos[j] - me for j in range(3) if j != i]
        if i == 0:  # adversary: landmarks, then other agents in world order
            obs = np.concatenate(entity_pos + other_pos)
        else:       # good agent: goal first
            obs = np.concatenate([self.landmarks[self.goal].p_pos - me] + entity_pos + other_pos)
        return obs.astype(np.float32)

    # -- rewards (A.5) --------------------------------------------------------
    def _dist_to_goal(self, i):
        d = self.p_pos[i] - self.landmarks[self.goal].p_pos
        return float(np.sqrt(np.sum(np.square(d))))

    def _reward(self, i):
        if i == 0:
            return -self._dist_to_goal(0)
        adv_rew = self._dist_to_goal(0)
        pos_rew = -min(self._dist_to_goal(1), self._dist_to_goal(2))
        return pos_rew + adv_rew

    # -- world step (A.2, A.4) ------------------------------------------------
    def _world_step(self):
        for i in range(3):
            a = int(self.current_actions[i]) % N_ACTIONS
            u = np.zeros(2)
            if a == 1:
                u[0] = -1.0
            if a == 2:
                u[0] = +1.0
            if a == 3:
                u[1] = -1.0
            if a == 4:
                u[1] = +1.0
            u *= SENSITIVITY
            if self.integrate_pos_first:
                self.p_pos[i] += self.p_vel[i] * DT
            self.p_vel[i] = self.p_vel[i] * (1 - DAMPING)
            self.p_vel[i] += (u / MASS) * DT
            if not self.integrate_pos_first:
                self.p_pos[i] += self.p_vel[i] * DT
        for i, a in enumerate(AGENT_ORDER):
            self.rewards[a] = self._reward(i)

    # -- AEC step / last (A.7) ------------------------------------------------
    def step(self, action):
        cur = self.agent_selection
        if self.terminations[cur] or self.truncations[cur]:
            # upstream _was_dead_step: remove the agent (never reached by the
            # reference, which breaks on the first truncation flag)
            self.agents.remove(cur)
            if self.agents:
                self._sel = self._sel % len(self.agents)
                self.agent_selection = self.agents[self._sel]
            return
        i = self._index[cur]
        nxt = (i + 1) % 3
        self._sel = nxt
        self.agent_selection = AGENT_ORDER[nxt]
        self.current_actions[i] = action
        if nxt == 0:
            self._world_step()
            self.steps += 1
            if self.steps >= self.max_cycles:
                for a in AGENT_ORDER:
                    self.truncations[a] = True
        else:
            for a in AGENT_ORDER:
                self.rewards[a] = 0.0
        self._cumulative_rewards[cur] = 0.0
        for a in AGENT_ORDER:
            self._cumulative_rewards[a] += self.rewards[a]

    def last(self, observe=True):
        a = self.agent_selection
        obs = self.observe(a) if observe else None
        return (obs, self._cumulative_rewards[a], self.terminations[a],
                self.truncations[a], self.infos[a])

    def close(self):
        pass

    def render(self):
        pass


def env(render_mode=None, **kwargs):
    """Factory with the upstream module-level name (``simple_adversary_v3.env``)."""
    return SimpleAdversaryEnv(render_mode=render_mode, **kwargs)


def draw_initial_states(n, seed=1870300):
    """``n`` flat initial-state records from the PCG64 stream the reference
    seeds at ``utils/game_logic_functions.py:54`` (one record per ``reset``;
    the first reset is the seeded one inside ``initialize_env``)."""
    e = SimpleAdversaryEnv()
    e.reset(seed=seed)
    for _ in range(n - 1):
        e.reset()
    return np.stack(e.init_state_log[:n])
