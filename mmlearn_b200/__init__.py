"""mmlearn_b200 -- B200-native (sm_100a) implementation of mmlearn's contrastive-loss hot path.

Public surface (mirrors the reference for this path only):

* :class:`ContrastiveLoss` (alias ``FusedContrastiveLoss``) -- drop-in for
  ``mmlearn.modules.losses.ContrastiveLoss``
* :class:`LossPairSpec`, :data:`Modalities` -- interface mirrors
* :func:`find_matching_indices` -- device version of the reference function
* :class:`RetrievalRecallAtK` -- drop-in for ``mmlearn.modules.metrics.RetrievalRecallAtK`` on the
  same similarity mainloop
* :class:`ZeroShotScores` -- the scoring step of ``mmlearn.tasks.ZeroShotClassification`` (scores
  and top-k ranks of queries against class embeddings) on that mainloop

Everything computes through ``libmmcl.so`` (``include/mmcl.h``); there is no CPU fallback.
"""

from . import _lib
from .loss import ContrastiveLoss, FusedContrastiveLoss, find_matching_indices, get_handle
from .modalities import LossPairSpec, Modalities
from .plan import PairPlan
from .retrieval import RetrievalRecallAtK
from .zero_shot import ZeroShotScores

__all__ = [
    "ContrastiveLoss",
    "FusedContrastiveLoss",
    "LossPairSpec",
    "Modalities",
    "PairPlan",
    "RetrievalRecallAtK",
    "ZeroShotScores",
    "find_matching_indices",
    "get_handle",
]
