"""Shim for eight_mile.pytorch.layers (TEST INFRASTRUCTURE ONLY): re-exports oracle/eight_mile_compat.py."""
from eight_mile_compat import *  # noqa: F401,F403
from eight_mile_compat import (  # noqa: F401
    pytorch_conv1d, pytorch_linear, PassThru, Conv1DSame, TransformerEncoderStack, Dense, MaxPool1D, MeanPool1D,
    TwoHeadConcat, SingleHeadReduction, BasicDualEncoderModel, sequence_mask, sequence_mask_mxlen, EmbeddingsStack,
    TransformerDecoderStack, WeightTieDense, subsequent_mask,
)
