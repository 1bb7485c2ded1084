"""tip_b200 -- B200-native (sm_100a) implementation of the TIP tri-graph encoder/decoder hot path.

    from tip_b200.layers import *        # drop-in for the reference's `from src.layers import *`
    from tip_b200.neg_sampling import typed_negative_sampling
    from tip_b200.utils import process_edges, to_bidirection, get_range_list

Device work goes through libtipb200.so (C ABI: include/tipb200.h); importing the package does
not need a GPU, calling any operator does.
"""
from ._lib import LIB_PATH, TipbError, lib  # noqa: F401

__version__ = "0.1.0"
