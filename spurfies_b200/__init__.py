"""spurfies_b200 -- B200-native (sm_100a) implementation of Spurfies' per-ray training/rendering hot path.

Only what that path needs lives here: ``csrc/`` (CUDA kernels + the C ABI, include/spurfies_b200.h),
and the host-side mirrors of the reference interfaces (``VoxelGrid`` = torch_knnquery.VoxelGrid,
``PointVolSDF`` = spurfies.model.pointneus_disent.PointVolSDF).  There is no CPU fallback: the
kernels are reached through ``libspurfies_b200.so`` and importing the ops without it raises.
"""
__version__ = "0.1.0"
