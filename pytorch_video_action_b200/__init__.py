"""pytorch_video_action_b200: B200-native MS-TCN hot path of mrqorib/pytorch-video-action.

Host-side mirror of the reference interface (networks.MultiStageModel, the CE criterion, the
argmax / segment-vote post-processing) over hand-written sm_100a kernels behind a C ABI
(include/mstcn_b200.h).  No CPU path, no PyTorch fallback.
"""
from .networks import MultiStageModel, SingleStageModel, DilatedResidualLayer  # noqa: F401
from .loss import FrameCrossEntropy, MsTcnLoss  # noqa: F401
from .postprocess import frame_argmax, segment_vote, ensemble_vote, label_runs, evaluate_video, ensemble_predict  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .graph import GraphedTrainStep  # noqa: F401
from .data import DeviceFeatureStore, RaggedBatchUploader  # noqa: F401

__all__ = ["MultiStageModel", "SingleStageModel", "DilatedResidualLayer", "FrameCrossEntropy", "MsTcnLoss", "frame_argmax",
           "segment_vote", "ensemble_vote", "label_runs", "evaluate_video", "ensemble_predict", "FusedAdam", "GraphedTrainStep", "DeviceFeatureStore", "RaggedBatchUploader"]
