r"""Rank-1 modifications of Cholesky factorisations on the B200.

Same surface as the reference's ``cholupdates.rank_1`` (``src/cholupdates/rank_1/__init__.py:39-49``):
``update``, ``downdate``, ``update_seeger``, ``downdate_seeger``; plus the in-place aliases and the
batched entry points.
"""

from ._update import update, update_in_place  # isort: skip
from ._downdate import downdate, downdate_in_place  # isort: skip

# Concrete update functions
from ._seeger import update_seeger  # isort: skip
from ._seeger import downdate_seeger  # isort: skip

from ._batched import update_batched, downdate_batched  # isort: skip

__all__ = [
    "update",
    "downdate",
    "update_in_place",
    "downdate_in_place",
    "update_seeger",
    "downdate_seeger",
    "update_batched",
    "downdate_batched",
]
