from coevonet_b200.utils.utils_pth_and_plots import *  # noqa: F401,F403  (drop-in shim)
