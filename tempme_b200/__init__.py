"""tempme_b200 -- B200 (sm_100a) implementation of TempME's motif hot path behind the reference's
Python API.  See DESIGN.md; the C ABI is include/tempme_b200.h."""
from ._lib import TempMEError, launch_count, lib  # noqa: F401
from .graph import NeighborFinder, class_hist_device, edge_identity_device, new_edge_info  # noqa: F401
from .null_model import RandEdgeSampler, degree_dict, get_null_distribution, load_data_shuffle, pre_processing, statistic  # noqa: F401
from .explainer import TempME, TimeEncode  # noqa: F401
from .pipeline import MotifPipeline  # noqa: F401
from .pack import build_pack, load_pack, save_pack  # noqa: F401

__all__ = ["NeighborFinder", "TempME", "TimeEncode", "MotifPipeline", "get_null_distribution", "RandEdgeSampler",
           "new_edge_info", "class_hist_device", "edge_identity_device", "launch_count", "build_pack", "save_pack", "load_pack"]
