from coevonet_b200.Atari.atari_agent import *  # noqa: F401,F403  (drop-in shim)
