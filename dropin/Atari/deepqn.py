from coevonet_b200.Atari.deepqn import *  # noqa: F401,F403  (drop-in shim)
