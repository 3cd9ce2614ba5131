"""nsc_b200 -- B200-native implementation of NSC's batched frame-wise codec pass.

Operator surface (same names as the reference's modules):
    nsc_b200.nn_core_operator, nsc_b200.lpc_utilities, nsc_b200.loss_terms_and_measures, nsc_b200.constants
Harness over the fused C-ABI entry points:
    nsc_b200.codec (CodecConfig, NeuralCodec, CMRL)
Everything computes in libnsc_b200.so (hand-written sm_100a CUDA); there is no CPU fallback.
"""
__version__ = "0.1.0"
