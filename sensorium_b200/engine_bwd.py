"""Backward plan of DwiseNeuro over the hand-written kernels (formulas: SURVEY.md §7.4).

``run_backward`` consumes the state saved by ``engine.run_forward(..., save=True)`` and the gradients of the
readout outputs and returns one gradient (or None) per parameter in ``mod.parameters()`` order.  Mice whose
output gradient is None keep ``grad=None`` exactly like the reference, where ``MicePoissonLoss`` skips absent
mice (losses.py:17) and torch AdamW then skips those tensors.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import call, gemm
from .engine import (BF16, F32, _P, _P_SDW, _empty, _fork, _join, _nullctx, _p_sdw, _p_tdw, _shadow, _side_streams,
                     _split_k, _stream)

_J_CORTEX = 32
_J_TDW = 9   # tdw_bwd_reduce CTAs per sample (kbench KB_JTDW sweep: 9 is best at every block size)


def _bn_bwd(part, P, NQ, q0, count, bn, grads, C, st, dev):
    bcoef = _empty((2, C), torch.float32, dev)
    dgamma = _empty((C,), torch.float32, dev)
    dbeta = _empty((C,), torch.float32, dev)
    call("dwn_bn_bwd_finalize", part, P, NQ, q0, float(count), dgamma, dbeta, bcoef, C, st)
    grads[bn.weight] = dgamma
    grads[bn.bias] = dbeta
    return bcoef


def _bn_bwd2(part, P, NQ, q0a, q0b, count, bn_a, bn_b, grads, C, st, dev):
    """Two BatchNorms over the same partial table in one launch (dwn_bn_bwd_finalize2)."""
    out = []
    for _ in range(2):
        out.append((_empty((C,), torch.float32, dev), _empty((C,), torch.float32, dev), _empty((2, C), torch.float32, dev)))
    (dga, dba, bca), (dgb, dbb, bcb) = out
    call("dwn_bn_bwd_finalize2", part, P, NQ, q0a, q0b, float(count), dga, dba, bca, dgb, dbb, bcb, C, st)
    grads[bn_a.weight], grads[bn_a.bias] = dga, dba
    grads[bn_b.weight], grads[bn_b.bias] = dgb, dbb
    return bca, bcb


_wg_side = {}


def _wgrad_stream(dev):
    """Side stream of the weight-gradient chains (conv_pw wgrad GEMM -> split-K reduce -> finalize, depth-wise weight
    finalizes): nothing downstream in backward depends on them, so they leave the critical path of the captured graph."""
    if dev.index not in _wg_side:
        _wg_side[dev.index] = torch.cuda.Stream(device=dev)
    return _wg_side[dev.index]


def run_backward(mod, sv, grad_outs: Sequence[Optional[torch.Tensor]]) -> List[Optional[torch.Tensor]]:
    cfg = mod.cfg
    bf = sv.mode == "bf16"
    adt = torch.bfloat16 if bf else torch.float32
    dcode = BF16 if bf else F32
    es = 2 if bf else 4
    dev = sv.x.device
    st = _stream(dev)
    B, T = sv.B, sv.T
    G = cfg["groups"]
    Mbt = B * T
    grads: Dict[torch.Tensor, torch.Tensor] = {}
    if not mod.training:
        raise NotImplementedError("sensorium_b200: backward is implemented for train mode (batch-stat BatchNorm)")

    dp = getattr(mod, "_dp", None)
    nsm = _lib.lib().dwn_sm_count()
    if dp is not None:
        dp.begin([g is not None for g in grad_outs], dev, mice=[r.m for r in sv.readouts])
        if getattr(dp, "sm_reserve", 0) > 0:
            _lib.lib().dwn_set_sm_budget(max(nsm - dp.sm_reserve, 32))
            nsm = _lib.lib().dwn_sm_count()
    PB = 4 * nsm  # rows of the per-CTA partial tables (= persistent CTAs per channel chunk) of the trunk kernels

    # weight-gradient side chain (at most one outstanding): `pending` keeps the tensors it reads alive until main has
    # waited for it, so the allocator cannot hand their memory to a later main-stream allocation
    from . import engine as _engine
    wside = None if _engine.SERIALIZE else _wgrad_stream(dev)
    pending: list = []
    side_done = [None]
    deferred: list = []   # data-parallel: parameter groups whose exchange waits for the side chain that produces them

    def join_wside():
        if side_done[0] is not None:
            torch.cuda.current_stream(dev).wait_event(side_done[0])
            side_done[0] = None
        pending.clear()
        while deferred:
            dp.reduce(grads, deferred.pop(0))

    def run_on_wside(fn, keep):
        """fn(stream_handle) on the side stream after everything main has enqueued so far."""
        if wside is None:
            fn(st)
            return
        join_wside()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        wside.wait_event(ev)
        with torch.cuda.stream(wside):
            fn(_stream(dev))
        done = torch.cuda.Event()
        done.record(wside)
        side_done[0] = done
        pending.extend(keep)

    # ---------------- readouts -------------------------------------------------------------------
    K = cfg["cortex_features"][-1]
    Kg = K // G
    live = [(r, g) for r, g in zip(sv.readouts, grad_outs) if g is not None]
    dX = _empty((Mbt, K), torch.float32, dev)
    if live:
        dxm = _empty((len(live), Mbt, K), torch.float32, dev)
        has_mask = live[0][0].mask is not None
        masks = torch.stack([r.mask.reshape(B, K) for r, _ in live]).contiguous() if has_mask else None
        # allocations on the main stream, then the independent per-mouse kernels round-robin on side streams
        jobs = []
        for j, (r, g) in enumerate(live):
            conv = mod.readouts[r.m].layer[1]
            half = r.half
            half_pad = ((half + 63) // 64) * 64
            wr = conv.weight
            jobs.append((r, g.contiguous(), conv, half, half_pad, _empty((G * half, Mbt), adt, dev),
                         _empty((Mbt, G * half_pad), adt, dev), _empty((G * half,), torch.float32, dev),
                         _empty((G * half, Kg, 1), torch.float32, dev), _shadow(wr) if bf else wr))
        side = _side_streams(dev, len(jobs))
        _fork(side, dev)
        for j, (r, g, conv, half, half_pad, dz_nm, dz_mn, db, dW, wq) in enumerate(jobs):
            with torch.cuda.stream(side[j % len(side)]) if side else _nullctx():
                sst = _stream(dev)
                call("dwn_readout_bwd_prep", r.pred, g, cfg["softplus_beta"], dz_nm, dz_mn, db, B, T, r.n_out, half,
                     half_pad, G, dcode, sst)
                gemm(sst, dtype=dcode, A=dz_nm, B=r.xt, lda=Mbt, ldb=Mbt, a_zstride=half * Mbt, b_zstride=Kg * Mbt,
                     a_zmode=1, b_zmode=1, M=half, N=Kg, K=Mbt, Z=G, D=dW, d_dtype=F32, ldd=Kg, d_zstride=half * Kg,
                     _tag="readout_wgrad", _bytes=G * half * Mbt * es + K * Mbt * es + G * half * Kg * 4)
                gemm(sst, dtype=dcode, A=dz_mn, B=wq, b_mn=1, lda=G * half_pad, ldb=Kg, a_zstride=half_pad,
                     b_zstride=half * Kg, a_zmode=1, b_zmode=1, M=Mbt, N=Kg, K=half, Z=G, D=dxm[j], d_dtype=F32, ldd=K,
                     d_zstride=Kg, _tag="readout_dgrad",
                     _bytes=Mbt * G * half_pad * es + G * half * Kg * es + Mbt * K * 4)
            grads[conv.weight] = dW
            grads[conv.bias] = db
        _join(side, dev)
        call("dwn_readout_dx_combine", dxm, masks, len(live), dX, Mbt, K, T, st)
    else:
        dX.zero_()
    if dp is not None:
        dp.reduce_readouts(grads, mice=[r.m for r in sv.readouts])
    early = getattr(mod, "_early_step", None)
    if early is not None and dp is None and wside is not None and live:
        # the readout gradients (95 % of the parameters) are final here: their AdamW + EMA update runs on a side stream
        # under the remaining ~16 ms of backward, whose kernels leave HBM bandwidth unused (MouseModel.train_step)
        early(grads, [p_ for r, _ in live for p_ in mod.readouts[r.m].layer[1].parameters()])

    # ---------------- cortex ---------------------------------------------------------------------
    dOut = dX
    for layer, c in zip(reversed(list(mod.cortex.layers)), reversed(sv.cortex)):
        I, O = c.I, c.O
        part = _empty((_J_CORTEX, 4, O), torch.float32, dev)
        call("dwn_cortex_bwd_reduce", dOut, c.Y, c.coef, c.dp, c.x, c.coef_sc, part, _J_CORTEX, Mbt, T, I, O, G, dcode, st)
        bcoef, bcoef_sc = _bn_bwd2(part, _J_CORTEX, 4, 0, 2, Mbt, layer.bn.bn, layer.bn_sc.bn, grads, O, st, dev)
        dY = _empty((Mbt, O), adt, dev)
        call("dwn_cortex_bwd_dy", dOut, c.Y, c.coef, bcoef, c.dp, dY, Mbt, T, O, G, dcode, st)
        dW = _empty((O, I // G, 1), torch.float32, dev)
        gemm(st, dtype=dcode, A=dY, B=c.xb if bf else c.x, a_mn=1, b_mn=1, lda=O, ldb=I, a_zstride=O // G,
             b_zstride=I // G, a_zmode=1, b_zmode=1, M=O // G, N=I // G, K=Mbt, Z=G, D=dW, d_dtype=F32, ldd=I // G,
             d_zstride=(O // G) * (I // G))
        grads[layer.conv.weight] = dW
        wc = layer.conv.weight
        dXc = _empty((Mbt, I), torch.float32, dev)
        gemm(st, dtype=dcode, A=dY, B=_shadow(wc) if bf else wc, b_mn=1, lda=O, ldb=I // G, a_zstride=O // G,
             b_zstride=(O // G) * (I // G), a_zmode=1, b_zmode=1, M=Mbt, N=I // G, K=O // G, Z=G, D=dXc, d_dtype=F32,
             ldd=I, d_zstride=I // G)
        dXn = _empty((Mbt, I), torch.float32, dev)
        call("dwn_cortex_in_bwd", dXc, dOut, c.x, c.coef_sc, bcoef_sc, dXn, Mbt, I, O, st)
        dOut = dXn
        if dp is not None:
            dp.reduce(grads, list(layer.parameters()))

    # ---------------- pool -----------------------------------------------------------------------
    HW, CL = sv.pool.HW, sv.pool.C
    dO = _empty((Mbt * HW, CL), torch.float32, dev)
    call("dwn_pool_bwd", dOut, dO, Mbt, HW, CL, st)

    # ---------------- inverted-residual blocks ---------------------------------------------------
    nb = len(sv.blocks)
    for i in range(nb - 1, -1, -1):
        b = sv.blocks[i]
        blk = mod.core.blocks[2 * i + 1]
        ci, co, mid, s, rd = b.ci, b.co, b.mid, b.s, b.rd
        Mi, Mo, Nsp = B * T * b.Hi * b.Wi, B * T * b.Ho * b.Wo, T * b.Ho * b.Wo
        part = _empty((PB, 4, co), torch.float32, dev)
        call("dwn_block_bwd_reduce", dO, b.Y, b.coef4, b.dp, b.X, b.coef_sc, part, PB, B, T, b.Ho, b.Wo, ci, co, s, b.Hi,
             b.Wi, dcode, st, _tag="block_bwd_reduce", _bytes=Mo * (co * (4 + es) + ci * 4))
        bcoef4, bcoef_sc = _bn_bwd2(part, PB, 4, 0, 2, Mo, blk.conv_pwl[1].bn, blk.bn_sc.bn, grads, co, st, dev)
        dY = _empty((Mo, co), adt, dev)
        call("dwn_block_bwd_dy", dO, b.Y, b.coef4, bcoef4, b.dp, dY, Mo, Nsp, co, dcode, st, _tag="block_bwd_dy",
             _bytes=Mo * co * (4 + 2 * es))
        # per-sample projection wgrad  Pp[b][mid][co] = a_b^T dY_b  -> dW_pwl and the SE gate gradient
        Pp = _empty((B, mid, co), torch.float32, dev)
        gemm(st, dtype=dcode, A=b.A, B=dY, a_mn=1, b_mn=1, lda=mid, ldb=co, a_zstride=Nsp * mid, b_zstride=Nsp * co,
             a_zmode=1, b_zmode=1, M=mid, N=co, K=Nsp, Z=B, D=Pp, d_dtype=F32, ldd=co, d_zstride=mid * co, _tag="pwl_wgrad",
             _bytes=Mo * (mid + co) * es + B * mid * co * 4)
        dpre2 = _empty((B, mid), torch.float32, dev)
        dhpre = _empty((B, rd), torch.float32, dev)
        dmean = _empty((B, mid), torch.float32, dev)
        se = blk.se
        dwpwl = torch.empty_like(blk.conv_pwl[0].weight)
        dw2, db2 = torch.empty_like(se.conv_expand.weight), torch.empty_like(se.conv_expand.bias)
        dw1, db1 = torch.empty_like(se.conv_reduce.weight), torch.empty_like(se.conv_reduce.bias)
        wt = blk.conv_pwl[0].weight.detach().reshape(co, mid).t().contiguous()
        call("dwn_se_bwd", Pp, wt, b.gate, b.hpre, b.mean, se.conv_reduce.weight,
             se.conv_expand.weight, dpre2, dhpre, dmean, dwpwl, dw2, db2, dw1, db1, B, mid, co, rd, st)
        grads[blk.conv_pwl[0].weight] = dwpwl
        grads[se.conv_expand.weight], grads[se.conv_expand.bias] = dw2, db2
        grads[se.conv_reduce.weight], grads[se.conv_reduce.bias] = dw1, db1
        # projection dgrad with the gated per-sample weights: da = dY Wb
        da = _empty((Mo, mid), adt, dev)
        gemm(st, dtype=dcode, A=dY, B=b.Wb, b_mn=1, lda=co, ldb=mid, a_zstride=Nsp * co, b_zstride=co * mid, a_zmode=1,
             b_zmode=1, M=Nsp, N=mid, K=co, Z=B, D=da, d_dtype=dcode, ldd=mid, d_zstride=Nsp * mid, _tag="pwl_dgrad",
             _bytes=(Mo * co + B * co * mid + Mo * mid) * es)
        # temporal dw backward
        jt = _J_TDW if 2 * nsm >= B * _J_TDW else max(1, 2 * nsm // B)   # one resident wave (2 CTAs per SM)
        part = _empty((B * jt, 2, mid), torch.float32, dev)
        call("dwn_tdw_bwd_reduce", da, b.Tm, b.coef3, dmean, Nsp, part, jt, B, mid, dcode, st, _tag="tdw_bwd_reduce",
             _bytes=2 * Mo * mid * es)
        bcoef3 = _bn_bwd(part, B * jt, 2, 0, Mo, blk.temp_covn_dw[1].bn, grads, mid, st, dev)
        ptdw = _p_tdw(Mo * mid) * nsm // 148
        part7 = _empty((ptdw, 7, mid), torch.float32, dev)
        call("dwn_tdw_bwd", da, b.Tm, b.S, b.coef3, bcoef3, b.coef2, blk.temp_covn_dw[0].weight, dmean, part7, ptdw, B, T,
             b.Ho * b.Wo, mid, dcode, st, _tag="tdw_bwd", _bytes=4 * Mo * mid * es)
        bcoef2 = _bn_bwd(part7, ptdw, 7, 0, Mo, blk.spat_covn_dw[1].bn, grads, mid, st, dev)
        dwt = torch.empty_like(blk.temp_covn_dw[0].weight)
        grads[blk.temp_covn_dw[0].weight] = dwt                 # filled by the side chain below
        # spatial dw backward (da now holds d s_hat)
        dE = _empty((Mi, mid), adt, dev)
        psdw = 42 if bf else _P_SDW     # TMA-staged backward: 42 workers per channel chunk are best on every block (KB_PS sweep)
        part11 = _empty((psdw, 11, mid), torch.float32, dev)
        call("dwn_sdw_bwd", da, b.S, b.E, b.coef2, bcoef2, b.coef1, blk.spat_covn_dw[0].weight, dE, part11, psdw, B * T,
             b.Hi, b.Wi, mid, s, dcode, st, _tag="sdw_bwd", _bytes=(2 * Mo + 2 * Mi) * mid * es)
        del da
        bcoef1 = _bn_bwd(part11, psdw, 11, 0, Mi, blk.conv_pw[1].bn, grads, mid, st, dev)
        dws = torch.empty_like(blk.spat_covn_dw[0].weight)
        grads[blk.spat_covn_dw[0].weight] = dws                 # filled by the side chain below
        wpw = blk.conv_pw[0].weight
        dXpw = _empty((Mi, ci), torch.float32, dev)
        tiles = math.ceil(mid / 128) * math.ceil(ci / 256)
        Zs = _split_k(Mi, tiles)
        rows = Mi // Zs
        wpart = _empty((Zs, mid, ci), torch.float32, dev)
        dwpw = torch.empty_like(wpw)
        colbias = None
        if bf and b.gram is not None:
            # BN1 backward folded into the GEMMs (dE holds G = dE_pre; E is never read): dwn_pw_algebra.cu
            wsh = _shadow(wpw)
            psum = _empty((mid, ci), torch.float32, dev)

            def wchain(sst, dE=dE, b=b, blk=blk, part7=part7, part11=part11, ptdw=ptdw, psdw=psdw, dwt=dwt, dws=dws,
                       wpart=wpart, psum=psum, dwpw=dwpw, wsh=wsh, bcoef1=bcoef1, rows=rows, Zs=Zs, mid=mid, ci=ci, Mi=Mi):
                # weight gradients of this block: nothing downstream in backward reads them
                gemm(sst, dtype=dcode, A=dE, B=b.Xb, a_mn=1, b_mn=1, lda=mid, ldb=ci, a_zstride=rows * mid,
                     b_zstride=rows * ci, a_zmode=1, b_zmode=1, M=mid, N=ci, K=rows, Z=Zs, D=wpart, d_dtype=F32, ldd=ci,
                     d_zstride=mid * ci, _tag="pw_wgrad", _bytes=Mi * (mid + ci) * es + Zs * mid * ci * 4)
                call("dwn_reduce_rows", wpart, Zs, mid * ci, psum, sst)
                call("dwn_pw_wgrad_finalize", psum, b.coef1, bcoef1, wsh, b.gram, b.sx, dwpw, mid, ci, sst)
                call("dwn_dw_wgrad_finalize", part7, ptdw, 7, 2, 5, dwt, mid, sst)
                call("dwn_dw_wgrad_finalize", part11, psdw, 11, 2, 9, dws, mid, sst)

            run_on_wside(wchain, [dE, part7, part11, wpart, psum, bcoef1, wsh])
            wprime = _empty((mid, ci), torch.bfloat16, dev)
            negq = _empty((ci, ci), torch.bfloat16, dev)
            colbias = _empty((ci,), torch.float32, dev)
            scratch = _empty((_lib.lib().dwn_pw_bwd_prep_scratch(mid, ci),), torch.float32, dev)
            call("dwn_pw_bwd_prep", b.coef1, bcoef1, wsh, wprime, negq, colbias, scratch, mid, ci, st)
            gemm(st, dtype=dcode, A=dE, B=wprime, b_mn=1, lda=mid, ldb=ci, M=Mi, N=ci, K=mid, Z=1, A2=b.Xb, B2=negq,
                 lda2=ci, ldb2=ci, K2=ci, D=dXpw, d_dtype=F32, ldd=ci, _tag="pw_dgrad",
                 _bytes=Mi * mid * es + mid * ci * es + Mi * ci * (4 + es))
        else:
            call("dwn_dw_wgrad_finalize", part7, ptdw, 7, 2, 5, dwt, mid, st)
            call("dwn_dw_wgrad_finalize", part11, psdw, 11, 2, 9, dws, mid, st)
            call("dwn_bn_bwd_apply", dE, b.E, b.coef1, bcoef1, Mi, mid, dcode, st, _tag="bn_bwd_apply",
                 _bytes=3 * Mi * mid * es)
            gemm(st, dtype=dcode, A=dE, B=_shadow(wpw) if bf else wpw, b_mn=1, lda=mid, ldb=ci, M=Mi, N=ci, K=mid, Z=1,
                 D=dXpw, d_dtype=F32, ldd=ci, _tag="pw_dgrad", _bytes=Mi * mid * es + mid * ci * es + Mi * ci * 4)
            gemm(st, dtype=dcode, A=dE, B=b.Xb if bf else b.X, a_mn=1, b_mn=1, lda=mid, ldb=ci, a_zstride=rows * mid,
                 b_zstride=rows * ci, a_zmode=1, b_zmode=1, M=mid, N=ci, K=rows, Z=Zs, D=wpart, d_dtype=F32, ldd=ci,
                 d_zstride=mid * ci, _tag="pw_wgrad", _bytes=Mi * (mid + ci) * es + Zs * mid * ci * 4)
            call("dwn_reduce_rows", wpart, Zs, mid * ci, dwpw, st)
        grads[wpw] = dwpw
        del dE
        stem_fused = i == 0 and mod.core.stem[0].weight.shape[1] == 5
        if stem_fused:
            # block 0: the gradient w.r.t. the stem output feeds the stem reductions directly (never stored)
            p_stem = _lib.lib().dwn_block_in_bwd_stem_rows(ci)
            stem_part = _empty((p_stem, 6, ci), torch.float32, dev)
            call("dwn_block_in_bwd_stem", dXpw, dO, b.X, b.coef_sc, bcoef_sc, colbias, sv.x, stem_part, p_stem, B, T,
                 b.Hi, b.Wi, ci, co, s, st, _tag="block_in_bwd", _bytes=Mi * ci * 8 + Mo * co * 4 + Mi * 20)
            dO = None
        else:
            dXin = _empty((Mi, ci), torch.float32, dev)
            call("dwn_block_in_bwd", dXpw, dO, b.X, b.coef_sc, bcoef_sc, colbias, dXin, B, T, b.Hi, b.Wi, ci, co, s, st,
                 _tag="block_in_bwd", _bytes=Mi * ci * 12 + Mo * co * 4)
            dO = dXin
        if dp is not None:
            # the exchange reads this block's weight gradients, which the side chain is still producing: it is issued at
            # the next join (one block later, same order on every rank)
            deferred.append(list(blk.parameters()))
            if wside is None or side_done[0] is None:
                join_wside()

    join_wside()
    if dp is not None and getattr(dp, "sm_reserve", 0) > 0:
        _lib.lib().dwn_set_sm_budget(0)
    # ---------------- stem -----------------------------------------------------------------------
    stem_conv, stem_bn = mod.core.stem[0], mod.core.stem[1].bn
    cin = stem_conv.weight.shape[1]
    C0 = stem_conv.weight.shape[0]
    dw = torch.empty_like(stem_conv.weight)
    dgam, dbet = torch.empty_like(stem_bn.weight), torch.empty_like(stem_bn.bias)
    if dO is None:
        call("dwn_stem_bwd_finalize", stem_part, p_stem, sv.stem.mom, stem_conv.weight, sv.stem.coef, dw, dgam, dbet, B, cin,
             sv.T * sv.H * sv.W, C0, st)
    else:
        part = _empty((_P, cin + 1, C0), torch.float32, dev)
        call("dwn_stem_bwd", dO, sv.x, part, _P, sv.stem.mom, stem_conv.weight, sv.stem.coef, dw, dgam, dbet, B, cin,
             sv.T * sv.H * sv.W, C0, st)
    grads[stem_conv.weight], grads[stem_bn.weight], grads[stem_bn.bias] = dw, dgam, dbet
    if dp is not None:
        dp.reduce(grads, [stem_conv.weight, stem_bn.weight, stem_bn.bias])
        dp.finish(dev)

    return [grads.get(p) if p.requires_grad else None for p in mod.parameters()]
