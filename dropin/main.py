from coevonet_b200.main import *  # noqa: F401,F403  (drop-in shim: same top-level module name as the reference)
