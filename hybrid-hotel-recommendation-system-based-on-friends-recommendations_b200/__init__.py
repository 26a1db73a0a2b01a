"""dcnr_b200 -- B200-native (sm_100a) implementation of the DCN-R hot path of
Krist-Marrakesh/Hybrid-Hotel-Recommendation-System-Based-on-Friends-Recommendations.

Public surface (drop-in for the reference's objects on this path):
  DCN_RecSys, CrossLayer, ResBlock      train.py:90-170 / main.py:61-127
  NearestNeighbors                      sklearn object built at main.py:268-269
  serving.rank_candidates               main.py:319-325
Everything computes through libdcnr_sm100a.so (include/dcnr.h); there is no CPU fallback.
"""
from . import _cabi
from .knn import NearestNeighbors, merge_shards
from .model import DCN_RecSys, CrossLayer, ResBlock, CrossLayerV2, CrossNetworkV2
from . import functional
from . import serving
from . import distributed
from . import training

__all__ = ["DCN_RecSys", "CrossLayer", "ResBlock", "CrossLayerV2", "CrossNetworkV2", "NearestNeighbors", "merge_shards", "functional", "serving",
           "distributed", "training", "library_path", "launch_count"]


def library_path() -> str:
    return _cabi.LIB_PATH


def launch_count(reset: bool = False) -> int:
    """Kernels launched by the library on this thread (bench.py's gpu_launches)."""
    return _cabi.launch_count(reset)
