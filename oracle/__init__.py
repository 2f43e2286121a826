"""CPU oracle for the CoEvoNet population-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import it, and there only as the checker / the timed CPU baseline.
The product path (``coevonet_b200``) never imports this package and fails
loudly when its CUDA library is missing.

Parity status
-------------
* Network forward, argmax, rollout attribution, GA/ES arithmetic, diversity,
  DeepQN forward: **pinned** against the reference's own Python code imported
  from ``/root/reference`` in the build container (``oracle/make_golden.py``
  generated ``tests/golden/*.npz``; the reference has no golden vectors or
  tests of its own).
* ``simple_adversary_v3`` physics / observations / rewards / AEC bookkeeping:
  **parity unpinned** -- the arithmetic lives in the third-party dependency
  ``pettingzoo`` (unpinned in the reference's ``requirements.txt:5``, Python
  3.9 => most plausibly 1.24.x) which is not vendored in the reference and not
  installable here (no network).  ``oracle/mpe_env.py`` restates the published
  upstream algorithm and is anchored on the reference's
  call sites (``utils/game_logic_functions.py:46,54,130,138,179,181,217``).
"""
