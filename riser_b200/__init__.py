"""riser_b200 -- B200-native implementation of the RISER read-classification hot
path (trim -> normalise -> 1-D CNN -> decision) behind the reference's
Kit / SignalProcessor / Model / SequencerControl call surface."""
from .preprocess import Kit, SignalProcessor, RaggedBatch      # noqa: F401
from .model import Model, decide, PREC_F16, PREC_F16_W2, PREC_F16_X3, PREC_F16_F8   # noqa: F401

from .pipeline import BatchedClassifier, FixedBatchPipeline     # noqa: F401,E402
from .control import SequencerControl                           # noqa: F401,E402
from .resnet import ResNetModel                                 # noqa: F401,E402

__all__ = ["Kit", "SignalProcessor", "RaggedBatch", "Model", "decide", "PREC_F16", "PREC_F16_W2", "PREC_F16_X3", "PREC_F16_F8",
           "BatchedClassifier", "FixedBatchPipeline", "SequencerControl",
           "ResNetModel"]
