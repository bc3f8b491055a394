"""partsbaseddetector_b200 -- B200-native (sm_100a) hot path of wg-perception/PartsBasedDetector:
HOG pyramid -> part-filter responses -> distance-transform tree DP -> backtrack, as hand-written CUDA kernels
behind a C-ABI (include/pbd_b200.h, libpbd_b200.so) and this thin host mirror of the reference interface."""
from .flatmodel import FlatModel, FlatPart  # noqa: F401
from .detector import Candidate, CandidateList, FileStorageModel, MatlabIOModel, Model, PartsBasedDetector, Dt2dPlan, dt2d, filterCandidatesByDepth  # noqa: F401
from ._lib import PbdError, SO_PATH, build  # noqa: F401

__all__ = ["Candidate", "FileStorageModel", "MatlabIOModel", "Model", "PartsBasedDetector", "PbdError", "FlatModel", "FlatPart", "dt2d", "Dt2dPlan", "filterCandidatesByDepth", "build", "SO_PATH"]
