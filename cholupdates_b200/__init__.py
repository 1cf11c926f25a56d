"""cholupdates_b200 -- B200-native (sm_100a) rank-1 Seeger Cholesky update / downdate.

Drop-in for one hot path of ``marvinpfoertner/cholupdates``: ``cholupdates.rank_1.update`` /
``downdate`` (reference ``src/cholupdates/rank_1/__init__.py:39-49``).  Same arguments, same
exceptions; NumPy arrays or torch CUDA tensors; hand-written CUDA behind a C ABI
(``include/cholupdates_b200.h``); no CPU fallback.
"""

from . import rank_1

__version__ = "0.1.0"

__all__ = ["rank_1"]
