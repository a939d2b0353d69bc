"""`stnf` drop-in package: the upstream ST-DADK module interface on the B200-native st_dadk_b200 kernels."""
__version__ = "0.1.0"
