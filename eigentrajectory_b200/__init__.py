"""eigentrajectory_b200 -- B200 (sm_100a) implementation of the EigenTrajectory descriptor hot path.

Same class and function names as the reference's ``EigenTrajectory`` package and
``utils.metrics`` so that a predictor plugin or script switches by changing its import:

    from eigentrajectory_b200 import EigenTrajectory, ETDescriptor, ETAnchor, TrajNorm, BatchKMeans
    from eigentrajectory_b200 import compute_batch_ade, compute_batch_fde

Everything numerical runs in ``libet_b200.so`` (hand-written CUDA behind a C ABI, see
``include/et_b200.h``); importing this package without the built library raises on first use.
"""
from . import dataloader, ops  # noqa: F401
from ._lib import ETLibraryError, load as load_library, launch_count  # noqa: F401
from .anchor import ETAnchor  # noqa: F401
from .descriptor import ETDescriptor  # noqa: F401
from .kmeans import BatchKMeans  # noqa: F401
from .metrics import (compute_batch_ade, compute_batch_ade_fde, compute_batch_col, compute_batch_fde,  # noqa: F401
                      compute_batch_metric, compute_batch_tcc)
from .model import EigenTrajectory  # noqa: F401
from .normalizer import TrajNorm  # noqa: F401
from .utils import DotDict  # noqa: F401

__all__ = ["EigenTrajectory", "ETDescriptor", "ETAnchor", "TrajNorm", "BatchKMeans", "compute_batch_ade",
           "compute_batch_fde", "compute_batch_ade_fde", "compute_batch_tcc", "compute_batch_col", "compute_batch_metric", "DotDict", "ops", "dataloader", "load_library", "launch_count",
           "ETLibraryError"]
