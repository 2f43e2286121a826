from coevonet_b200.utils.utils_policies import *  # noqa: F401,F403  (drop-in shim)
