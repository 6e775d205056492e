"""Shared shapes: configs/true_batch_001.py:21-39 of the reference."""
TRUE_BATCH_KW = dict(in_channels=5, core_features=(64, 64, 64, 64, 128, 128, 128, 256, 256),
                     spatial_strides=(2, 1, 1, 1, 2, 1, 1, 2, 1), spatial_kernel=3, temporal_kernel=5,
                     expansion_ratio=7, se_reduce_ratio=32, cortex_features=(1024, 2048, 4096), groups=2,
                     softplus_beta=0.07, drop_rate=0.4, drop_path_rate=0.1)
TINY_KW = dict(core_features=(16, 16, 32), spatial_strides=(2, 1, 2), expansion_ratio=4, se_reduce_ratio=8,
               cortex_features=(64, 128), groups=2, drop_path_rate=0.3)
TINY_OUTS = (37, 64, 129)
