from coevonet_b200.MPE.fcnetwork import *  # noqa: F401,F403  (drop-in shim)
