"""B200-native photometric-alignment hot path behind the reference's estimator API.

    from dense_visual_odometry_b200 import get_dvo, RGBDCameraModel, Se3
    dvo = get_dvo("robust-dvo", camera_model, Se3.identity(), levels=4)
    T = dvo.step(color_bgr_u8, depth_u16)

Everything numeric runs in libdvo_b200.so (hand-written sm_100a CUDA, C ABI in include/dvo_b200.h);
importing this package does not need a GPU, using it does.
"""
from .camera_model import RGBDCameraModel
from .lie import Se3, So3, pose_to_qt
from .estimator import (PairBatchAligner, RobustDVOB200, SequenceAligner, get_dvo, make_config, robust_dvo_factory,
                        stats_to_numpy)
from ._cabi import DvoError
from . import dataset
from .sharding import chain_poses, gather_poses, sequence_shard_range, shard_range

__all__ = ["RGBDCameraModel", "Se3", "So3", "pose_to_qt", "PairBatchAligner", "RobustDVOB200", "SequenceAligner", "get_dvo",
           "make_config", "robust_dvo_factory", "stats_to_numpy", "DvoError", "chain_poses", "gather_poses",
           "sequence_shard_range", "shard_range", "dataset"]
