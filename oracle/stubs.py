"""Import shims that let the UNMODIFIED reference run in the build container.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``/root/reference`` only
exists in the build container; nothing on the GPU box may call
:func:`import_reference`.  Used by ``oracle/make_golden.py`` (golden-vector
generation) and by the ``reference``-marked CPU tests, which skip when the
tree is absent.

Three third-party modules the reference imports are absent from this image
(SURVEY.md section 8c): ``supersuit`` (four wrapper names, Atari only),
``matplotlib.pyplot`` (plots) and ``pettingzoo.mpe.simple_adversary_v3``
(replaced by ``oracle.mpe_env``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("COEVONET_REFERENCE_ROOT", "/root/reference")


class _Anything:
    """Callable / attribute sink used for the plotting stub."""

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return self

    def __iter__(self):
        return iter(())


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    """Register the stub modules (idempotent)."""
    from . import mpe_env

    if "supersuit" not in sys.modules:
        ident = lambda env, *a, **k: env  # noqa: E731
        _module("supersuit", frame_stack_v1=ident, resize_v1=ident,
                frame_skip_v0=ident, agent_indicator_v0=ident)
    if "matplotlib" not in sys.modules:
        plt = types.ModuleType("matplotlib.pyplot")
        sink = _Anything()

        def _getattr(name):
            if name.startswith("__") and name.endswith("__"):
                raise AttributeError(name)
            return sink

        plt.__getattr__ = _getattr
        sys.modules["matplotlib.pyplot"] = plt
        _module("matplotlib", pyplot=plt)
    if "pettingzoo" not in sys.modules:
        pz = _module("pettingzoo")
        mpe = _module("pettingzoo.mpe")
        sys.modules["pettingzoo.mpe.simple_adversary_v3"] = mpe_env
        pz.mpe = mpe
        mpe.simple_adversary_v3 = mpe_env


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "genetic_algorithm.py"))


def import_reference():
    """Return a namespace with the reference's modules imported unmodified."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    ns = types.SimpleNamespace()
    for name in ("agent", "MPE.fcnetwork", "MPE.mpe_agent", "Atari.deepqn",
                 "utils.game_logic_functions", "genetic_algorithm",
                 "evolutionary_strategy"):
        setattr(ns, name.replace(".", "_"), importlib.import_module(name))
    return ns


class RefArgs:
    """The reference's mutable ``Args`` bag (``main.py:95-142``) with the
    ``train_GA.sh`` / ``train_ES.sh`` defaults, for driving it from tests."""

    def __init__(self, **kw):
        self.algorithm = "GA"
        self.generations = 2
        self.population = 20
        self.hof_size = 3
        self.game = "simple_adversary_v3"
        self.mutation_power_agent_0 = 0.005
        self.mutation_power_agent_1 = 0.05
        self.mutation_power_adversary = 0.05
        self.learning_rate = 0.1
        self.max_timesteps_per_episode = 400
        self.max_evaluation_steps = 400
        self.elites_number = 5
        self.adaptive = True
        self.max_mutation_power = 0.7
        self.min_mutation_power = 0.0001
        self.fitness_sharing = True
        self.early_stopping = False
        self.patience = 300
        self.min_delta = 0.1
        self.debug = False
        self.train = True
        self.test = False
        self.render = False
        self.env_mode = "AEC"
        self.precision = "float32"
        self.save = False
        self.play_against_yourself = False
        self.average_window = 50
        self.__dict__.update(kw)
