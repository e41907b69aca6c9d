"""ragmeup_b200 — B200-native dense-retrieval hot path of RAGMeUp (embed -> top-k -> rerank).

Host code is Python (PyTorch tensors at the boundary); all arithmetic runs in hand-written
sm_100a CUDA kernels behind the C ABI declared in ``include/ragmeup_b200.h``.
"""
__version__ = "0.1.0"
