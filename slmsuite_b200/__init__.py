"""
slmsuite_b200 -- B200-native GS / WGS hologram optimisation (drop-in for the
``Hologram`` / ``SpotHologram`` ``optimize()`` path of slmsuite.holography.algorithms).

Python host classes -> ctypes -> ``libslmgs.so`` (hand-written sm_100a CUDA, C ABI in
``include/slmgs.h``).  No PyTorch / CuPy on the hot path; see DESIGN.md.
"""

from .hologram import ALGORITHM_DEFAULTS, ALGORITHM_INDEX, FEEDBACK_OPTIONS, Hologram
from .spots import SpotHologram
from .batch import HologramBatch, optimize_sharded, shard_bounds
from .multiplane import MultiplaneHologram
from .camera import SimulatedCamera
from .compressed import CompressedSpotHologram, ShardedCompressedSpotHologram

__all__ = ["Hologram", "SpotHologram", "HologramBatch", "MultiplaneHologram", "SimulatedCamera", "CompressedSpotHologram", "ShardedCompressedSpotHologram", "optimize_sharded", "shard_bounds", "ALGORITHM_DEFAULTS", "ALGORITHM_INDEX", "FEEDBACK_OPTIONS"]
__version__ = "0.1.0"
