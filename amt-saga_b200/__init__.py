"""amt-saga_b200: B200-native STFT / CQT / generative-subtractive feature path
of AMT-SAGA behind the reference's `audio_complete` interface.

Python here is host glue only (plan construction, tensor ownership, the lazy
container); all arithmetic on the path runs in libsaga_b200.so (hand-written
sm_100a CUDA behind the C ABI in include/saga_b200.h).  There is no CPU
fallback: importing the compute modules without the built library raises.
"""
__version__ = "0.1.0"
