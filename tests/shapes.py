"""Shared shapes: configs/true_batch_001.py:21-39 of the reference."""
TRUE_BATCH_KW = dict(in_channels=5, core_features=(64, 64, 64, 64, 128, 128, 128, 256, 256),
                     spatial_strides=(2, 1, 1, 1, 2, 1, 1, 2, 1), spatial_kernel=3, temporal_kernel=5,
                     expansion_ratio=7, se_reduce_ratio=32, cortex_features=(1024, 2048, 4096), groups=2,
                     softplus_beta=0.07, drop_rate=0.4, drop_path_rate=0.1)
TINY_KW = dict(core_features=(16, 16, 32), spatial_strides=(2, 1, 2), expansion_ratio=4, se_reduce_ratio=8,
               cortex_features=(64, 128), groups=2, drop_path_rate=0.3)
TINY_OUTS = (37, 64, 129)


import torch  # noqa: E402


def corr_step_outputs():
    """Three validation batches for three mice (ragged neuron counts); mouse 1 never has a sample, some samples of the
    others are masked by a zero weight.  Shared by the golden generator and the tests (seeded, regenerated there)."""
    g = torch.Generator().manual_seed(11)
    outs_n = (37, 5, 130)
    steps = []
    for it in range(3):
        B = 4 + it
        w = torch.zeros(B, 3)
        ids = torch.randint(0, 2, (B,), generator=g) * 2          # mouse 0 or 2
        w[torch.arange(B), ids] = torch.rand(B, generator=g) + 0.1
        if it == 1:
            w[0] = 0.0                                            # a sample that belongs to no mouse
        preds = [torch.rand(B, n, 16, generator=g) * 4 for n in outs_n]
        targets = [((preds[m] * 0.7 + torch.randn(B, n, 16, generator=g)).relu() * 3) * (w[:, m] != 0)[:, None, None].float()
                   for m, n in enumerate(outs_n)]
        steps.append({"prediction": preds, "target": (targets, w)})
    return steps
