"""lp-gnn hot path, B200-native: bipartite constraint<->variable message passing (GCN_FC),
basis-status head, knowledge masking and basis selection as hand-written sm_100a CUDA kernels
behind the reference's own Python surface (arch.py / dataset.py / val.py / train.py /
scripts/pred_basis.py).  The kernels live in ``csrc/`` and are reached through the C-ABI
declared in ``include/lpgnn.h`` (``liblpgnn.so``, loaded with ctypes by ``_lib.py``).

There is no CPU fallback: every op raises if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
