"""B200-native Chamfer / kNN hot path of geometric_adv behind the reference's operator API.

    from geometric_adv_b200 import nn_distance, knn_point, group_point, chamfer_3DDist

The arithmetic lives in libga_b200.so (hand-written CUDA for sm_100a, C ABI in
include/ga_b200.h); this package is the thin PyTorch shim the reference's attack,
defense and evaluation scripts call.
"""
from .ops import (GA_MODE_CPU_EXACT, GA_MODE_GPU_REF, chamfer_3DDist, chamfer_3DFunction, chamfer_all_pairs,
                  chamfer_loss_terms, chamfer_per_cloud, group_point, knn_dists, knn_point, launch_count, nn_distance,
                  nn_distance_grad, select_top_k, set_default_mode, set_pruning)

from . import attack, defense, sharding  # noqa: E402,F401  (host loops either side of the hot path)

__version__ = "0.1.0"
