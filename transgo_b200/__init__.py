"""transgo_b200 -- B200-native batched self-play engine behind Transgo's environment / MCTS surface.

The compute lives in ``libtransgo_b200.so`` (hand-written sm_100a CUDA, C ABI in ``include/transgo_b200.h``);
this package is the thin PyTorch/ctypes host side that mirrors ``GoEnv/environment.py`` and the ``WP_MCTS``
entry point of ``self_play.py``.  There is no CPU fallback: importing the kernels without the built library,
or calling them without a CUDA device, raises.
"""
__version__ = "0.1.0"
