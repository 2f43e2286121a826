from coevonet_b200.MPE.mpe_agent import *  # noqa: F401,F403  (drop-in shim)
