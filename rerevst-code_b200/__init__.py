"""B200-native ReReVST per-frame stylization path.

Mirrors the reference's Python module API for the hot path only:
  style_network_global.TransformerNet   test/style_network_global.py:454-501
  framework.Stylization                 test/framework.py:56-118
  loss_networks.warp / TemporalLoss     train/loss_networks.py:20-111
All tensor arithmetic runs in csrc/librerevst_b200.so (hand-written sm_100a kernels, C ABI in
include/rerevst_b200.h).  There is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
