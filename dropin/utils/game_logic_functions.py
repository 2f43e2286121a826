from coevonet_b200.utils.game_logic_functions import *  # noqa: F401,F403  (drop-in shim)
from coevonet_b200.utils.game_logic_functions import _device  # noqa: F401
