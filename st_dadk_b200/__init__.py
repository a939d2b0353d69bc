"""st_dadk_b200 -- B200-native (sm_100a) hot path of ST-DADK: multi-resolution basis embedding of
(x, y, t) fused with the MLP forward/backward, behind the reference's `stnf` module interface.

Everything numerical runs in libstdadk.so (hand-written CUDA, C ABI in include/stdadk.h); PyTorch
provides device memory, streams and torch.distributed.  There is no CPU fallback."""
__version__ = "0.1.0"
