"""gedepth_b200 - B200-native (sm_100a) implementation of the GEDepth ground-embedding +
DepthFormer hot path behind the reference's mmcv-style registry surface."""
__version__ = "0.1.0"
