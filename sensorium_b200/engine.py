"""Kernel orchestration for DwiseNeuro on B200: forward / backward plans over libdwn_b200.so.

Data layout: every activation is channels-last ``[B][T][H][W][C]`` == row-major ``[M][C]``.
 * trunk (block inputs/outputs) is fp32 (+ a bf16 copy used as tcgen05 A operand in bf16 mode) — this
   mirrors the autocast dtype map of the reference (the residual trunk is fp32);
 * branch intermediates E_raw / S_raw / Tm_raw (pre-BatchNorm conv outputs) are stored in the pipeline
   dtype (bf16 or fp32); BatchNorm + SiLU are applied by the *consumer* on load, batch statistics are
   produced by the *producer* as deterministic per-CTA partial sums and finalised by a tiny kernel.

Reference call graph: DwiseNeuro.forward (/root/reference/src/models/dwiseneuro.py:397-405).
"""
from __future__ import annotations

import math
import weakref
from functools import partial as _partial
from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from ._lib import call, gemm

F32, BF16 = 0, 1
BN_MOM, BN_EPS = 0.1, 1e-5
_P = 592          # partial-stat rows for kernels whose grid.y may be 1 (4 CTAs / SM)
_P_SDW = 148      # spatial dw kernels have >= 3 channel chunks in grid.y
_P_MOM = 296
_J_SE = 8           # se_pool CTAs per sample: 256 CTAs = one balanced wave (16 -> 512 CTAs = 1.15 waves at 3 CTAs/SM: 4.7 vs 5.9 TB/s)


def _p_tdw(nelem: int) -> int:
    """Persistent CTAs per channel chunk of the temporal depth-wise kernels: 592 for the large early blocks, 148 where a
    launch moves < 150 M elements (blocks 4-8 at C2; measured with tests/gpu_checks/kbench.py, KB_P sweep: the late
    blocks gain 5-15 % with fewer, longer-lived CTAs, the early blocks lose 35 %)."""
    return _P if nelem > 150_000_000 else 148


def _p_sdw(nelem_in: int) -> int:
    """Same for the spatial depth-wise kernels (KB_PS sweep): 42 workers per channel chunk unless the input tile stream is
    the 940 M-element block-0 tensor (42 vs 148 workers: -2 ... -15 % on blocks 4-8, neutral on blocks 1-3)."""
    return _P_SDW if nelem_in > 600_000_000 else 42

_plans = weakref.WeakKeyDictionary()


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


def pe_tables(pe_mod, C: int, T: int, H: int, W: int, dev):
    """Separable lookup tables of PositionalEncoding3d (dwiseneuro.py:147-182).

    PE[c,t,h,w] = pe_t[t][c] + pe_h[h][c] + pe_w[w][c] where exactly one term is non-zero: channels
    [0,ch) encode T, [ch,2ch) encode H, [2ch,3ch) encode W (truncated to C); inside an axis run the first
    half is sin(pos*f_k), the second half cos(pos*f_k)."""
    cache = _plans.setdefault(pe_mod, {})
    key = (C, T, H, W, str(dev))
    if key in cache:
        return cache[key]
    inv_freq = pe_mod.inv_freq.to(device=dev, dtype=torch.float32)
    ch = pe_mod.channels
    tabs = []
    for axis, n in enumerate((T, H, W)):
        pos = torch.arange(n, device=dev).type(inv_freq.type())
        ang = torch.einsum("i,j->ij", inv_freq, pos)           # (ch/2, n)
        emb = torch.cat((ang.sin(), ang.cos()), dim=0)         # (ch, n)
        tab = torch.zeros(n, 3 * ch, dtype=torch.float32, device=dev)
        tab[:, axis * ch:(axis + 1) * ch] = emb.t()
        tabs.append(tab[:, :C].contiguous())
    cache[key] = tuple(tabs)
    return cache[key]


def _shadow(p: torch.Tensor) -> torch.Tensor:
    """bf16 shadow of an fp32 GEMM weight, refreshed when the parameter version / storage changes.

    The cache lives on the parameter object itself (tensors cannot key a WeakKeyDictionary: ``==`` is
    element-wise)."""
    ent = getattr(p, "_dwn_shadow", None)
    if ent is not None and ent[0] == p._version and ent[2] == p.data_ptr() and ent[1].device == p.device:
        return ent[1]
    sh = torch.empty(p.shape, dtype=torch.bfloat16, device=p.device)
    call("dwn_cast_bf16", p.detach(), sh, p.numel(), _stream(p.device))
    p._dwn_shadow = (p._version, sh, p.data_ptr())
    return sh


def set_shadow(p: torch.Tensor, sh: torch.Tensor) -> None:
    """Used by the fused optimizer, which writes the bf16 shadow itself."""
    p._dwn_shadow = (p._version, sh, p.data_ptr())


# fp32 mode, eval (val_step / predict, argus_models.py:73-99): run the GEMMs on the tensor cores with fp32 accuracy — every
# fp32 operand is split into three bf16 planes (dwn_split3) and the six significant plane products are accumulated in the
# fp32 TMEM tile (dwn_gemm split = 3).  Training in fp32 mode keeps the FFMA GEMM.
FP32_TENSOR_CORES = True


def _planes(t: torch.Tensor, st) -> torch.Tensor:
    n = t.numel()
    pl = torch.empty((3, (n + 7) // 8 * 8), dtype=torch.bfloat16, device=t.device)
    call("dwn_split3", t, pl, n, st)
    if pl.shape[1] != n:
        raise RuntimeError("split operand size must be a multiple of 8")
    return pl


def _param_planes(p: torch.Tensor, st) -> torch.Tensor:
    """Split planes of a weight, cached like the bf16 shadows (version + storage + raw-write generation)."""
    key = (p._version, p.data_ptr(), _GEN[0], str(p.device))
    ent = getattr(p, "_dwn_planes", None)
    if ent is not None and ent[0] == key:
        return ent[1]
    pl = _planes(p.detach(), st)
    p._dwn_planes = (key, pl)
    return pl


def gemm_f32(st, training: bool, a_param=None, b_param=None, a_planes=None, **kw):
    """fp32-mode GEMM: tensor cores with split operands in eval mode when the shapes allow TMA (multiples of 8),
    FFMA otherwise.  ``a_param`` / ``b_param``: the parameter an operand IS (its split planes are cached);
    ``a_planes``: A already is a plane triple written by its producer."""
    A, B = kw["A"], kw["B"]
    if a_planes is not None:
        pb = _planes(B, st)
        return gemm(st, **dict(kw, A=a_planes, B=pb, dtype=BF16, split=3, a_pstride=a_planes.shape[1],
                               b_pstride=pb.shape[1]))
    ok = (FP32_TENSOR_CORES and not training and kw.get("A2") is None and kw["K"] % 8 == 0 and kw["lda"] % 8 == 0
          and kw["ldb"] % 8 == 0 and kw.get("a_zstride", 0) % 8 == 0 and kw.get("b_zstride", 0) % 8 == 0
          and A.numel() % 8 == 0 and B.numel() % 8 == 0 and A.is_contiguous() and B.is_contiguous())
    if not ok:
        return gemm(st, **kw)
    pa = _param_planes(a_param, st) if a_param is not None else _planes(A, st)
    pb = _param_planes(b_param, st) if b_param is not None else _planes(B, st)
    kw = dict(kw, A=pa, B=pb, dtype=BF16, split=3, a_pstride=pa.shape[1], b_pstride=pb.shape[1])
    return gemm(st, **kw)


_side = {}
# bench.py's per-kernel accounting sets this: every launch goes to the current stream, so CUDA-event pairs around single
# launches never overlap each other (side-stream work would otherwise be counted twice)
SERIALIZE = False


class _nullctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def _side_streams(dev, njobs, n=3):
    """Side streams for independent per-mouse launches (readout GEMMs)."""
    if njobs < 2 or SERIALIZE:
        return []
    key = (dev.index, n)
    if key not in _side:
        _side[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _side[key]


def _fork(side, dev):
    if side:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        for s in side:
            s.wait_event(ev)


def _join(side, dev):
    main = torch.cuda.current_stream(dev)
    for s in side:
        ev = torch.cuda.Event()
        ev.record(s)
        main.wait_event(ev)


_GEN = [0]


def bump_generation() -> None:
    """Called by everything that writes parameters / buffers through raw pointers (train-mode forward: running
    statistics; FusedAdamW.step; ModelEma.update; the data-parallel broadcast): torch's version counters do not see
    those writes, so caches derived from the weights (eval BatchNorm tables) are keyed on this counter as well."""
    _GEN[0] += 1


_FROZEN = [False]  # set by run_forward for modules flagged ``_dwn_frozen`` (teacher / predictor models: nothing writes
#                    their weights through raw pointers, so raw-write generations do not invalidate their tables)


def _eval_coef(bn, C, Cp, st, dev):
    """Eval-mode BatchNorm folded to per-channel (scale, shift) once per set of weights: the table is cached on the module
    and rebuilt only when a parameter / running statistic changes (version counters + storage + bump_generation)."""
    ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = (C, Cp, str(dev), -1 if _FROZEN[0] else _GEN[0]) + tuple((t._version, t.data_ptr()) for t in ts)
    ent = getattr(bn, "_dwn_eval_coef", None)
    if ent is not None and ent[0] == key:
        return ent[1]
    coef = _empty((4, C), torch.float32, dev)
    call("dwn_bn_finalize", None, 0, 1.0, bn.weight, bn.bias, bn.running_mean, bn.running_var, None,
         BN_MOM, BN_EPS, 0, coef, C, Cp, 2, st)
    bn._dwn_eval_coef = (key, coef)
    return coef


def _bn_coef(bn, partial, P, count, C, Cp, training, st, dev, NQ=2, out=None):
    if not training:
        return _eval_coef(bn, C, Cp, st, dev)
    coef = _empty((4, C), torch.float32, dev) if out is None else out
    call("dwn_bn_finalize", partial, P, float(count), bn.weight, bn.bias, bn.running_mean, bn.running_var,
         bn.num_batches_tracked, BN_MOM, BN_EPS, 1, coef, C, Cp, NQ, st)
    return coef


def _split_k(rows: int, tiles: int) -> int:
    z = 1
    while z * 2 * tiles <= 160 and rows % (z * 2) == 0 and rows // (z * 2) >= 256:
        z *= 2
    return z


def _gram(xb, M, ci, sx, st, dev, out=None):
    """Gx = X^T X (ci x ci, fp32) of the bf16 block input via a split-K (MN,MN) tcgen05 GEMM; returns (Gx, centred
    second moment Gx/M - mu mu^T) with mu = sx / M."""
    # short accumulation chains: the tensor core adds every K=16 slice into the fp32 TMEM accumulator with truncation,
    # a bias that grows with the chain length; 2048 rows per split keep the Gram-derived statistics within 1e-4 of the
    # direct ones at M = 2.1 M (tests/test_parity_fullsize_gpu.py::test_gram_batchnorm_statistics_full_rows)
    zs = 1
    while M % (zs * 2) == 0 and M // (zs * 2) >= 2048:
        zs *= 2
    rows = M // zs
    part = _empty((zs, ci, ci), torch.float32, dev)
    gemm(st, dtype=BF16, A=xb, B=xb, a_mn=1, b_mn=1, lda=ci, ldb=ci, a_zstride=rows * ci, b_zstride=rows * ci,
         a_zmode=1, b_zmode=1, M=ci, N=ci, K=rows, Z=zs, D=part, d_dtype=F32, ldd=ci, d_zstride=ci * ci, _tag="gram",
         _bytes=M * ci * 2)
    g = _empty((ci, ci), torch.float32, dev) if out is None else out
    cg = _empty((ci, ci), torch.float32, dev)
    call("dwn_gram_finalize", part, zs, ci, sx, float(M), g, cg, st)
    return g, cg


_stats_side: Dict[int, "torch.cuda.Stream"] = {}


def _stats_stream(dev):
    """Side stream for the Gram-matrix BatchNorm statistics of conv_pw (they depend only on the block input and
    run next to the expansion GEMM)."""
    if dev.index not in _stats_side:
        _stats_side[dev.index] = torch.cuda.Stream(device=dev)
    return _stats_side[dev.index]


def _colstats(x, M, ld, C, dcode, st, dev):
    part = _empty((_P, 2, C), torch.float32, dev)
    call("dwn_colstats", x, M, ld, C, part, _P, dcode, st, _tag="colstats", _bytes=M * C * (2 if dcode == BF16 else 4))
    return part


def _drop_mask(shape, keep, dtype, dev, rng_dev=None):
    # same torch calls (shape, dtype, order) as drop_path (dwiseneuro.py:46-54) so RNG streams agree;
    # rng_dev="cpu" draws from the CPU generator (parity tests against CPU-generated golden vectors)
    m = torch.empty(shape, dtype=dtype, device=rng_dev or dev).bernoulli_(keep)
    if keep > 0.0:
        m.div_(keep)
    return m.float().reshape(shape[0]).contiguous().to(dev)


def run_forward(mod, x: torch.Tensor, index: Optional[int], mode: str, training: bool, save: bool):
    """Returns (list of predictions [B, n_m, T] fp32, saved-state or None)."""
    cfg = mod.cfg
    dev = x.device
    st = _stream(dev)
    bf = mode == "bf16"
    adt = torch.bfloat16 if bf else torch.float32
    dcode = BF16 if bf else F32
    es = 2 if bf else 4
    x = x.detach().contiguous().float()
    B, Cin, T, H, W = x.shape
    feats = cfg["core_features"]
    strides = cfg["spatial_strides"]
    nb = len(feats)
    er = cfg["expansion_ratio"]
    G = cfg["groups"]
    if cfg["spatial_kernel"] != 3 or cfg["temporal_kernel"] != 5:
        raise NotImplementedError("sensorium_b200 kernels are specialised for spatial_kernel=3, temporal_kernel=5")
    sv = SimpleNamespace(blocks=[], cortex=[], readouts=[], mode=mode, B=B, T=T, H=H, W=W, x=x) if save else None
    if training:
        bump_generation()                              # running statistics are about to change
    _FROZEN[0] = bool(getattr(mod, "_dwn_frozen", False)) and not training
    rng_dev = getattr(mod, "_rng_device", None)        # test hooks: where / in which dtype the masks are drawn
    mdt = getattr(mod, "_mask_dtype", None) or adt

    # drop-path masks of the blocks and the cortex layers, drawn up front in the reference's call order (drop_path,
    # dwiseneuro.py:46-54 - the RNG stream only sees the order of the calls, not when they run) on a side stream: three tiny
    # torch kernels per mask no longer sit between a block's projection and its residual epilogue
    dp_blocks: List[Optional[torch.Tensor]] = [None] * nb
    dp_cortex: List[Optional[torch.Tensor]] = [None] * len(mod.cortex.layers)
    mask_side = []
    if training:
        mask_side = [_stats_stream(dev)] if (rng_dev is None and dev.type == "cuda" and not SERIALIZE) else []
        _fork(mask_side, dev)
        with torch.cuda.stream(mask_side[0]) if mask_side else _nullctx():
            for i in range(nb):
                blk_ = mod.core.blocks[2 * i + 1]
                if blk_.drop_path_rate > 0.0:
                    dp_blocks[i] = _drop_mask((B, 1, 1, 1, 1), 1.0 - blk_.drop_path_rate, mdt, dev, rng_dev)
            for li, layer_ in enumerate(mod.cortex.layers):
                if layer_.drop_path_rate > 0.0:
                    dp_cortex[li] = _drop_mask((B, 1, 1), 1.0 - layer_.drop_path_rate, mdt, dev, rng_dev)
    masks_joined = not mask_side
    if mask_side:
        for t_ in dp_blocks + dp_cortex:
            if t_ is not None:
                t_.record_stream(torch.cuda.current_stream(dev))  # allocated under the side stream, consumed on this one

    # ---------------- stem (dwiseneuro.py:306-309) + PE of block 0 -------------------------------
    stem_conv, stem_bn = mod.core.stem[0], mod.core.stem[1].bn
    C0 = feats[0]
    M0 = B * T * H * W
    coef0 = _empty((4, C0), torch.float32, dev) if training else None
    mom = None
    if training:
        nm = Cin + Cin * (Cin + 1) // 2
        momp = _empty((_P_MOM, nm), torch.float64, dev)
        mom = _empty((nm,), torch.float64, dev)
        call("dwn_input_moments", x, B, Cin, T * H * W, momp, _P_MOM, mom, st)
        call("dwn_stem_coef", mom, Cin, float(M0), stem_conv.weight, stem_bn.weight, stem_bn.bias,
             stem_bn.running_mean, stem_bn.running_var, stem_bn.num_batches_tracked, BN_MOM, BN_EPS, coef0, C0, st)
    else:
        coef0 = _eval_coef(stem_bn, C0, 0, st, dev)
    pe = pe_tables(mod.core.blocks[0], C0, T, H, W, dev)
    X = _empty((M0, C0), torch.float32, dev)
    Xb = _empty((M0, C0), torch.bfloat16, dev) if bf else None
    sc_part = _empty((_P, 3, C0), torch.float32, dev) if training else None
    call("dwn_stem_fwd", x, stem_conv.weight, coef0, pe[0], pe[1], pe[2], X, Xb, sc_part, _P, strides[0], B, Cin, T, H,
         W, C0, st, _tag="stem_fwd", _bytes=M0 * (Cin * 4 + C0 * (4 + (2 if bf else 0))))
    if save:
        sv.stem = SimpleNamespace(coef=coef0, mom=mom)

    # ---------------- inverted-residual blocks (dwiseneuro.py:136-144) ---------------------------
    Hi, Wi = H, W
    for i in range(nb):
        blk = mod.core.blocks[2 * i + 1]
        ci = feats[i]
        co = feats[i + 1] if i < nb - 1 else feats[-1]
        s = strides[i]
        mid = ci * er
        # conv (k=3, pad=1, stride s) and interpolate_shortcut (dwiseneuro.py:127-129) both produce ceil(size / s)
        Ho, Wo = -(-Hi // s), -(-Wi // s)
        Mi, Mo, Nsp = B * T * Hi * Wi, B * T * Ho * Wo, T * Ho * Wo
        # 1. point-wise expansion (tcgen05 GEMM / SIMT fp32)
        wpw = blk.conv_pw[0].weight
        E = _empty((Mi, mid), adt, dev)
        wsh = _shadow(wpw) if bf else wpw
        gram = sx = coef_sc_early = None
        stats_side = []
        if training and bf:
            # BatchNorm statistics of E = X W^T from the Gram matrix of X (no pass over E), see dwn_pw_algebra.cu.
            # They depend only on X: outputs are allocated here (main stream), the small kernels run on a side
            # stream next to the expansion GEMM and are joined before the spatial depth-wise kernel reads coef1.
            gram = _empty((ci, ci), torch.float32, dev)
            sx = _empty((ci,), torch.float32, dev)
            coef1 = _empty((4, mid), torch.float32, dev)
            coef_sc_early = _empty((4, co), torch.float32, dev)
            bn1 = blk.conv_pw[1].bn
            stats_side = [] if SERIALIZE else [_stats_stream(dev)]
            _fork(stats_side, dev)
            with torch.cuda.stream(stats_side[0]) if stats_side else _nullctx():
                sst = _stream(dev)
                # the shortcut BatchNorm's table depends only on the block input as well: off the critical path
                _bn_coef(blk.bn_sc.bn, sc_part, _P, Mo, co, ci, training, sst, dev, NQ=3, out=coef_sc_early)
                call("dwn_partial_colsum", sc_part, _P, 3, 2, ci, sx, sst)
                _, cgram = _gram(Xb, Mi, ci, sx, sst, dev, out=gram)
                call("dwn_pw_stats", cgram, sx, wsh, float(Mi), bn1.weight, bn1.bias, bn1.running_mean,
                     bn1.running_var, bn1.num_batches_tracked, BN_MOM, BN_EPS, coef1, mid, ci, sst)
        (gemm if bf else _partial(gemm_f32, training=training, b_param=wpw))(
            st, dtype=dcode, A=Xb if bf else X, B=wsh, lda=ci, ldb=ci, M=Mi, N=mid, K=ci, Z=1,
            D=E, d_dtype=dcode, ldd=mid, _tag="pw_fwd", _bytes=(Mi * ci + mid * ci + Mi * mid) * es)
        if gram is not None:
            _join(stats_side, dev)
        else:
            coef1 = _bn_coef(blk.conv_pw[1].bn, _colstats(E, Mi, mid, mid, dcode, st, dev) if training else None, _P,
                             Mi, mid, 0, training, st, dev)
        # 2. spatial depth-wise (BN1+SiLU on load)
        S = _empty((Mo, mid), adt, dev)
        part = _empty((_P, 2, mid), torch.float32, dev) if training else None
        psdw = _p_sdw(Mi * mid) if bf else _P_SDW   # the fp32 (generic) kernel is latency-bound: it wants every CTA it can get
        call("dwn_sdw_fwd", E, coef1, blk.spat_covn_dw[0].weight, S, part, psdw, B * T, Hi, Wi, mid, s, dcode, st,
             _tag="sdw_fwd", _bytes=(Mi + Mo) * mid * es)
        coef2 = _bn_coef(blk.spat_covn_dw[1].bn, part, psdw, Mo, mid, 0, training, st, dev)
        # 3. temporal depth-wise (BN2+SiLU on load)
        Tm = _empty((Mo, mid), adt, dev)
        part = _empty((_P, 2, mid), torch.float32, dev) if training else None
        ptdw = _p_tdw(Mo * mid)
        call("dwn_tdw_fwd", S, coef2, blk.temp_covn_dw[0].weight, Tm, part, ptdw, B, T, Ho * Wo, mid, dcode, st,
             _tag="tdw_fwd", _bytes=2 * Mo * mid * es)
        coef3 = _bn_coef(blk.temp_covn_dw[1].bn, part, ptdw, Mo, mid, 0, training, st, dev)
        # 4. squeeze-excite: a = SiLU(BN3(Tm)), gate folded into per-sample projection weights
        rd = blk.se.conv_reduce.weight.shape[0]
        pool_part = _empty((B, _J_SE, mid), torch.float32, dev)
        # fp32 eval: the squeeze kernel writes the three bf16 planes of a directly (operand of the split GEMM)
        a_planes = (not bf) and (not training) and FP32_TENSOR_CORES and mid % 8 == 0 and (Mo * mid) % 8 == 0
        if a_planes:
            A = _empty((3, Mo * mid), torch.bfloat16, dev)
            call("dwn_se_pool", Tm, coef3, A, pool_part, _J_SE, B, Nsp, mid, 2, st, _tag="se_pool",
                 _bytes=Mo * mid * 10)
        else:
            A = _empty((Mo, mid), adt, dev)
            call("dwn_se_pool", Tm, coef3, A, pool_part, _J_SE, B, Nsp, mid, dcode, st, _tag="se_pool",
                 _bytes=2 * Mo * mid * es)
        mean = _empty((B, mid), torch.float32, dev)
        hpre = _empty((B, rd), torch.float32, dev)
        gate = _empty((B, mid), torch.float32, dev)
        call("dwn_se_mlp", pool_part, _J_SE, Nsp, blk.se.conv_reduce.weight, blk.se.conv_reduce.bias,
             blk.se.conv_expand.weight, blk.se.conv_expand.bias, mean, hpre, gate, B, mid, rd, st)
        Wb = _empty((B, co, mid), adt, dev)
        call("dwn_fold_gate", blk.conv_pwl[0].weight, gate, Wb, B, co, mid, dcode, st)
        # 5. point-wise linear projection, batched over samples (B operand = gated weights of the sample)
        Y = _empty((Mo, co), adt, dev)
        (gemm if bf else _partial(gemm_f32, training=training, a_planes=A if a_planes else None))(
            st, dtype=dcode, A=A, B=Wb, lda=mid, ldb=mid, a_zstride=Nsp * mid, b_zstride=co * mid, a_zmode=1,
            b_zmode=1, M=Nsp, N=co, K=mid, Z=B, D=Y, d_dtype=dcode, ldd=co, d_zstride=Nsp * co, _tag="pwl_fwd",
            _bytes=(Mo * mid + B * co * mid + Mo * co) * es)
        coef4 = _bn_coef(blk.conv_pwl[1].bn, _colstats(Y, Mo, co, co, dcode, st, dev) if training else None, _P, Mo, co,
                         0, training, st, dev)
        coef_sc = coef_sc_early if coef_sc_early is not None else _bn_coef(blk.bn_sc.bn, sc_part, _P, Mo, co, ci, training,
                                                                              st, dev, NQ=3)
        # 6. residual epilogue (+ drop-path, + PE of the next block, + stats of the next shortcut)
        if not masks_joined:
            _join(mask_side, dev)
            masks_joined = True
        dp = dp_blocks[i]
        last = i == nb - 1
        pe = (None, None, None) if last else pe_tables(mod.core.blocks[2 * i + 2], co, T, Ho, Wo, dev)
        Xn = _empty((Mo, co), torch.float32, dev)
        Xnb = _empty((Mo, co), torch.bfloat16, dev) if (bf and not last) else None
        nsc_part = _empty((_P, 3, co), torch.float32, dev) if (training and not last) else None
        call("dwn_block_out", Y, coef4, dp, X, coef_sc, pe[0], pe[1], pe[2], Xn, Xnb, nsc_part, _P,
             1 if last else strides[i + 1], B, T, Ho, Wo, ci, co, s, Hi, Wi, dcode, st, _tag="block_out",
             _bytes=Mo * (co * es + ci * 4 + co * (4 + (2 if bf else 0))))
        if save:
            sv.blocks.append(SimpleNamespace(X=X, Xb=Xb, E=E, S=S, Tm=Tm, A=A, Y=Y, Wb=Wb, coef1=coef1, coef2=coef2,
                                             coef3=coef3, coef4=coef4, coef_sc=coef_sc, gate=gate, hpre=hpre,
                                             mean=mean, dp=dp, ci=ci, co=co, mid=mid, s=s, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo,
                                             rd=rd, gram=gram, sx=sx))
        X, Xb, sc_part, Hi, Wi = Xn, Xnb, nsc_part, Ho, Wo

    # ---------------- pool (dwiseneuro.py:374,400) -----------------------------------------------
    Mbt = B * T
    CL = feats[-1]
    cx = _empty((Mbt, CL), torch.float32, dev)
    cxb = _empty((Mbt, CL), torch.bfloat16, dev) if bf else None
    call("dwn_pool_hw", X, cx, cxb, Mbt, Hi * Wi, CL, st)
    if save:
        sv.pool = SimpleNamespace(HW=Hi * Wi, C=CL)

    # ---------------- cortex (dwiseneuro.py:195-263) ---------------------------------------------
    for li, layer in enumerate(mod.cortex.layers):
        I, O = layer.in_features, layer.out_features
        wc = layer.conv.weight
        Yc = _empty((Mbt, O), adt, dev)
        (gemm if bf else _partial(gemm_f32, training=training, b_param=wc))(
            st, dtype=dcode, A=cxb if bf else cx, B=_shadow(wc) if bf else wc, lda=I, ldb=I // G, a_zstride=I // G,
            b_zstride=(O // G) * (I // G), a_zmode=1, b_zmode=1, M=Mbt, N=O // G, K=I // G, Z=G, D=Yc, d_dtype=dcode,
            ldd=O, d_zstride=O // G)
        coef = _bn_coef(layer.bn.bn, _colstats(Yc, Mbt, O, O, dcode, st, dev) if training else None, _P, Mbt, O, 0,
                        training, st, dev)
        coef_sc = _bn_coef(layer.bn_sc.bn, _colstats(cx, Mbt, I, I, F32, st, dev) if training else None, _P, Mbt, O, I,
                           training, st, dev)
        if not masks_joined:
            _join(mask_side, dev)
            masks_joined = True
        dp = dp_cortex[li]
        out = _empty((Mbt, O), torch.float32, dev)
        outb = _empty((Mbt, O), torch.bfloat16, dev) if bf else None
        call("dwn_cortex_out", Yc, coef, dp, cx, coef_sc, out, outb, Mbt, T, I, O, G, dcode, st)
        if save:
            sv.cortex.append(SimpleNamespace(x=cx, xb=cxb, Y=Yc, coef=coef, coef_sc=coef_sc, dp=dp, I=I, O=O))
        cx, cxb = out, outb

    # ---------------- readouts (dwiseneuro.py:266-287) -------------------------------------------
    K = cfg["cortex_features"][-1]
    Kg = K // G
    outs = cfg["readout_outputs"]
    mice = range(len(outs)) if index is None else [index]
    preds: List[torch.Tensor] = []
    p_drop = cfg["drop_rate"]
    # phase 1 (main stream): RNG masks in the reference's order and every allocation
    jobs = []
    for m in mice:
        n_out = outs[m]
        mask = None
        if training and p_drop > 0.0:
            mask = torch.empty((B, K, 1), dtype=torch.float32, device=rng_dev or dev).bernoulli_(1.0 - p_drop).div_(
                1.0 - p_drop).to(dev)
        if mask is None and not save:
            xm, xt = (cxb if bf else cx), None
            prep = False
        else:
            xm = _empty((Mbt, K), adt, dev)
            xt = _empty((K, Mbt), adt, dev) if save else None
            prep = True
        conv = mod.readouts[m].layer[1]
        wr = conv.weight
        jobs.append((m, n_out, math.ceil(n_out / G), conv, _shadow(wr) if bf else wr, mask, xm, xt, prep,
                     _empty((B, n_out, T), torch.float32, dev)))
    # phase 2: the per-mouse GEMMs are independent and each fills well under half of the SMs (~62 CTAs), so they
    # are issued round-robin on side streams and overlap on the device
    side = _side_streams(dev, len(jobs))
    _fork(side, dev)
    for j, (m, n_out, half, conv, wq, mask, xm, xt, prep, pred) in enumerate(jobs):
        with torch.cuda.stream(side[j % len(side)]) if side else _nullctx():
            sst = _stream(dev)
            if prep:
                call("dwn_readout_prep", cx, mask, xm, xt, Mbt, K, T, dcode, sst)
            (gemm if bf else _partial(gemm_f32, training=training, a_param=conv.weight))(
                sst, dtype=dcode, A=wq, B=xm, lda=Kg, ldb=K, a_zstride=half * Kg, b_zstride=Kg, a_zmode=1, b_zmode=1,
                M=half, N=Mbt, K=Kg, Z=G, epi=1, D=pred, bias=conv.bias, beta=cfg["softplus_beta"], Tn=T,
                n_out_total=n_out, row_offset_per_z=half, n_limit=Mbt, _tag="readout_fwd",
                _bytes=G * half * Kg * es + Mbt * K * es + B * n_out * T * 4)
        preds.append(pred)
        if save:
            sv.readouts.append(SimpleNamespace(m=m, mask=mask, xm=xm, xt=xt, pred=pred, n_out=n_out, half=half))
    _join(side, dev)
    if save:
        sv.cx = cx
        sv.index = index
    return preds, sv


def forward(mod, x: torch.Tensor, index: Optional[int]):
    mode = mod._mode()
    training = mod.training
    params = [p for p in mod.parameters()]
    need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if need_grad:
        from .autograd import DwiseNeuroFn
        outs = DwiseNeuroFn.apply(mod, index, mode, x, *params)
        outs = list(outs)
    else:
        with torch.no_grad():
            outs, _ = run_forward(mod, x, index, mode, training, save=False)
    return outs if index is None else outs[0]
