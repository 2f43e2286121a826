"""coevonet_b200 -- B200-native (sm_100a) population-evaluation hot path of CoEvoNet.

Host code mirrors the reference's Python call surface (``genetic_algorithm``,
``evolutionary_strategy``, ``agent``, ``utils.game_logic_functions``, ``MPE``,
``Atari``); the bodies call hand-written CUDA kernels through the C ABI in
``include/coevonet_b200.h`` (``coevonet_b200/csrc``).  There is no CPU fallback:
compute entry points raise when ``libcoevonet_b200.so`` is missing.
"""
__version__ = "0.1.0"
