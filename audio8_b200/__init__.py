"""audio8_b200 — B200-native (sm_100a) implementation of mead-ml/audio8's wav2vec2 training hot path.

Public surface mirrors the reference: `audio8_b200.wav2vec2` (create_model, create_acoustic_model,
create_loss, module classes) and `audio8_b200.ctc` (CTCLoss).  All arithmetic runs in hand-written CUDA
behind the C ABI in include/audio8_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
