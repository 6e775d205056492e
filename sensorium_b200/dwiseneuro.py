"""DwiseNeuro — drop-in ``nn.Module`` surface over the B200 kernel engine.

Mirrors the constructor, ``forward(x, index=None)`` signature and the *exact* parameter / buffer tree
(names, shapes, registration order — including the ``spat_covn_dw`` / ``temp_covn_dw`` spelling) of
``/root/reference/src/models/dwiseneuro.py:343-405`` so that ``state_dict()``, ``load_state_dict()``,
``copy.deepcopy`` (ModelEma, ema.py:40), ``init_weights`` (utils.py:46-63) and ``torch.optim`` work
unchanged.  The sub-modules are only parameter holders: all arithmetic runs in
``sensorium_b200.engine`` through the C ABI of ``libdwn_b200.so``.  There is no CPU / eager fallback.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import torch
from torch import nn


class _Holder(nn.Module):
    """A bare container: children / parameters are attached by the builder functions below."""

    def forward(self, *args, **kwargs):  # pragma: no cover - never called
        raise RuntimeError("sensorium_b200 sub-modules are parameter holders; call DwiseNeuro.forward")


def _bn_act(num_features: int, bn_cls) -> nn.Module:
    h = _Holder()
    h.bn = bn_cls(num_features)
    return h


def _pos_enc(channels: int) -> nn.Module:
    # buffer layout of PositionalEncoding3d (dwiseneuro.py:148-157)
    h = _Holder()
    h.orig_channels = channels
    ch = int(math.ceil(channels / 6) * 2)
    if ch % 2:
        ch += 1
    h.channels = ch
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))
    h.register_buffer("inv_freq", inv_freq)
    h.register_buffer("cached_encoding", None, persistent=False)
    return h


def _inverted_residual(cin: int, cout: int, ks: int, kt: int, stride: int, er: int, se_rr: int,
                       drop_path_rate: float) -> nn.Module:
    # parameter tree of InvertedResidual3d (dwiseneuro.py:70-123)
    h = _Holder()
    mid = cin * er
    h.spatial_stride = stride
    h.out_features = cout
    h.drop_path_rate = float(drop_path_rate)
    h.conv_pw = nn.Sequential(nn.Conv3d(cin, mid, (1, 1, 1), bias=False), _bn_act(mid, nn.BatchNorm3d))
    h.spat_covn_dw = nn.Sequential(
        nn.Conv3d(mid, mid, (1, ks, ks), stride=(1, stride, stride), padding=(0, ks // 2, ks // 2), groups=mid,
                  bias=False),
        _bn_act(mid, nn.BatchNorm3d))
    h.temp_covn_dw = nn.Sequential(
        nn.Conv3d(mid, mid, (kt, 1, 1), stride=(1, 1, 1), padding=(kt // 2, 0, 0), groups=mid, bias=False),
        _bn_act(mid, nn.BatchNorm3d))
    se = _Holder()
    rd = mid // se_rr
    se.conv_reduce = nn.Conv3d(mid, rd, (1, 1, 1), bias=True)
    se.conv_expand = nn.Conv3d(rd, mid, (1, 1, 1), bias=True)
    h.se = se
    h.conv_pwl = nn.Sequential(nn.Conv3d(mid, cout, (1, 1, 1), bias=False), _bn_act(cout, nn.BatchNorm3d))
    h.bn_sc = _bn_act(cout, nn.BatchNorm3d)
    return h


def _shuffle_layer(cin: int, cout: int, groups: int, drop_path_rate: float) -> nn.Module:
    # parameter tree of ShuffleLayer (dwiseneuro.py:195-210)
    h = _Holder()
    h.in_features, h.out_features, h.groups = cin, cout, groups
    h.drop_path_rate = float(drop_path_rate)
    h.conv = nn.Conv1d(cin, cout, (1,), groups=groups, bias=False)
    h.bn = _bn_act(cout, nn.BatchNorm1d)
    h.bn_sc = _bn_act(cout, nn.BatchNorm1d)
    return h


def _readout(cin: int, cout: int, groups: int, beta: float, drop_rate: float) -> nn.Module:
    # parameter tree of Readout (dwiseneuro.py:266-281)
    h = _Holder()
    h.out_features = cout
    h.layer = nn.Sequential(
        nn.Dropout1d(p=drop_rate),
        nn.Conv1d(cin, math.ceil(cout / groups) * groups, (1,), groups=groups, bias=True))
    return h


class DwiseNeuro(nn.Module):
    def __init__(self,
                 readout_outputs: Sequence[int],
                 in_channels: int = 5,
                 core_features: Sequence[int] = (64, 64, 64, 64, 128, 128, 128, 256, 256),
                 spatial_strides: Sequence[int] = (2, 1, 1, 1, 2, 1, 1, 2, 1),
                 spatial_kernel: int = 3,
                 temporal_kernel: int = 5,
                 expansion_ratio: int = 6,
                 se_reduce_ratio: int = 32,
                 cortex_features: Sequence[int] = (1024, 2048, 4096),
                 groups: int = 2,
                 softplus_beta: float = 0.07,
                 drop_rate: float = 0.4,
                 drop_path_rate: float = 0.1):
        super().__init__()
        core_features = tuple(core_features)
        spatial_strides = tuple(spatial_strides)
        cortex_features = tuple(cortex_features)
        num_blocks = len(core_features)
        assert num_blocks and num_blocks == len(spatial_strides)  # dwiseneuro.py:304
        self.cfg = dict(
            readout_outputs=tuple(int(v) for v in readout_outputs), in_channels=in_channels,
            core_features=core_features, spatial_strides=spatial_strides, spatial_kernel=spatial_kernel,
            temporal_kernel=temporal_kernel, expansion_ratio=expansion_ratio, se_reduce_ratio=se_reduce_ratio,
            cortex_features=cortex_features, groups=groups, softplus_beta=softplus_beta, drop_rate=drop_rate,
            drop_path_rate=drop_path_rate)
        # "auto": bf16 pipeline under torch.autocast (train_step, argus_models.py:50), fp32 otherwise
        # (val_step / predict run without autocast, argus_models.py:73-99).  May be forced to "bf16"/"fp32".
        self.precision = "auto"

        core = _Holder()
        c0 = core_features[0]
        core.stem = nn.Sequential(nn.Conv3d(in_channels, c0, (1, 1, 1), bias=False), _bn_act(c0, nn.BatchNorm3d))
        blocks = []
        nxt = c0
        for i in range(num_blocks):
            cin = core_features[i]
            if i < num_blocks - 1:
                nxt = core_features[i + 1]
            blocks += [_pos_enc(cin),
                       _inverted_residual(cin, nxt, spatial_kernel, temporal_kernel, spatial_strides[i],
                                          expansion_ratio, se_reduce_ratio, drop_path_rate * i / num_blocks)]
        core.blocks = nn.Sequential(*blocks)
        self.core = core
        self.pool = nn.AdaptiveAvgPool3d((None, 1, 1))
        cortex = _Holder()
        cortex.layers = nn.Sequential()
        prev = core_features[-1]
        for f in cortex_features:
            cortex.layers.append(_shuffle_layer(prev, f, groups, drop_path_rate))
            prev = f
        self.cortex = cortex
        self.readouts = nn.ModuleList(
            [_readout(cortex_features[-1], n, groups, softplus_beta, drop_rate) for n in readout_outputs])

    # --------------------------------------------------------------------------------------------
    def _mode(self) -> str:
        if self.precision in ("bf16", "fp32"):
            return self.precision
        return "bf16" if torch.is_autocast_enabled() else "fp32"

    def forward(self, x: torch.Tensor, index: Optional[int] = None):
        from . import engine
        if x.dim() != 5:
            raise RuntimeError("The input tensor has to be 5D")  # dwiseneuro.py:185-186
        if not x.is_cuda:
            raise RuntimeError("sensorium_b200.DwiseNeuro runs on CUDA (sm_100a) only: no CPU fallback")
        assert x.shape[1] == self.cfg["in_channels"]
        return engine.forward(self, x, index)
