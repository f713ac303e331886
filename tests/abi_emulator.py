"""TEST INFRASTRUCTURE — a numpy executor for the *head* entry points of the C ABI
(``include/vlnce_b200.h``), plugged into ``vlnce_b200.lib.dry_run(executor=...)``.

Purpose: the host side of the package is launch sequencing over raw pointers, leading
dimensions and column offsets.  With this executor the CPU test-suite runs those launch plans
on host memory and compares the result with the oracle, so pointer arithmetic, buffer reuse and
the hand-derived backward are checked numerically without a GPU (GPU time is scarce; the
``-m gpu`` tests then only have to establish that each CUDA kernel matches its contract).

It lives under ``tests/`` and is never imported by ``vlnce_b200``: the product has no CPU path
(``lib.require_cuda`` / a missing library raise).  The conv-trunk entry points are deliberately
NOT emulated — tests that need trunk features substitute the oracle's trunk output.
Every function mirrors the kernel of the same name (semantics documented in the header).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

F32, I64, I32 = np.float32, np.int64, np.int32


def _flat(ptr, n, dtype=F32):
    if ptr is None or ptr == 0:
        return None
    ct = {F32: ctypes.c_float, I64: ctypes.c_longlong, I32: ctypes.c_int, np.uint8: ctypes.c_ubyte}[dtype]
    return np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ct)), shape=(int(n),))


def _mat(ptr, rows, cols, ld, col_stride=1, dtype=F32):
    """(rows, cols) strided view: element [r, c] at ptr + r*ld + c*col_stride."""
    if ptr is None or ptr == 0:
        return None
    span = (rows - 1) * abs(ld) + (cols - 1) * abs(col_stride) + 1
    base = _flat(ptr, span, dtype)
    it = base.itemsize
    return np.lib.stride_tricks.as_strided(base, shape=(rows, cols), strides=(ld * it, col_stride * it))


def _ten3(ptr, R, P, D, ld_r, ld_p):
    span = (R - 1) * ld_r + (P - 1) * ld_p + D
    base = _flat(ptr, span)
    return np.lib.stride_tricks.as_strided(base, shape=(R, P, D), strides=(ld_r * 4, ld_p * 4, 4))


def _sig(x):
    return 1.0 / (1.0 + np.exp(-x))


class Emulator:
    def __init__(self):
        self.count = 0

    def __call__(self, name, args):
        fn = getattr(self, name[len("vlnce_"):], None)
        if fn is None:
            raise NotImplementedError(f"abi_emulator: {name} is not emulated (conv-trunk entry point?)")
        self.count += 1
        with np.errstate(all="ignore"):  # uninitialised workspace beyond a row's length is read but never used
            fn(*args)

    # ---- dense algebra ---------------------------------------------------------------
    def sgemm(self, A, lda_m, lda_k, B, ldb_k, ldb_n, C, ldc, bias, M, N, K, beta, act, ws, wsb, st):
        a = _mat(A, M, K, lda_m, lda_k).astype(np.float64)
        b = _mat(B, K, N, ldb_k, ldb_n).astype(np.float64)
        c = _mat(C, M, N, ldc)
        y = a @ b
        if bias:
            y = y + _flat(bias, N).astype(np.float64)[None, :]
        if beta != 0.0:
            y = y + beta * c.astype(np.float64)
        if act == 1:
            y = np.maximum(y, 0.0)
        elif act == 2:
            y = np.tanh(y)
        c[...] = y.astype(F32)

    def sgemm_grouped(self, descs, count, ws, wsb, st):
        import ctypes
        import struct

        raw = ctypes.string_at(descs, count * 128)
        for i in range(count):
            (A, lda_m, lda_k, B, ldb_k, ldb_n, C, ldc, bias, M, N, K, beta, act, _, _) = struct.unpack_from(
                "<12qdq2q", raw, i * 128)
            if B == 0:  # column sums of A over k
                a = _mat(A, M, K, lda_m, lda_k).astype(np.float64)
                c = _mat(C, M, N, ldc)
                y = np.repeat(a.sum(1, keepdims=True), N, axis=1)
                if beta != 0.0:
                    y = y + beta * c.astype(np.float64)
                c[...] = y.astype(F32)
            else:
                self.sgemm(A, lda_m, lda_k, B, ldb_k, ldb_n, C, ldc, bias, M, N, K, beta, act, ws, wsb, st)

    def colsum(self, A, lda, out, M, N, accumulate, ws, wsb, st):
        s = _mat(A, M, N, lda).astype(np.float64).sum(0)
        o = _flat(out, N)
        o[...] = (s + (o if accumulate else 0.0)).astype(F32)

    # ---- forward ---------------------------------------------------------------------
    def embedding_fwd(self, ids, table, out, rows, dim, ldo, vocab, st):
        i = _flat(ids, rows, I64).copy()
        i[(i < 0) | (i >= vocab)] = 0
        _mat(out, rows, dim, ldo)[...] = _mat(table, vocab, dim, dim)[i]

    def _pa_index(self, prev, masks, rows, table_rows):
        i = ((_flat(prev, rows, I64).astype(F32) + 1.0) * _flat(masks, rows)).astype(I64)
        i[(i < 0) | (i >= table_rows)] = 0
        return i

    def prev_action_embed(self, prev, masks, table, out, ldo, out2, ldo2, rows, dim, table_rows, st):
        v = _mat(table, table_rows, dim, dim)[self._pa_index(prev, masks, rows, table_rows)]
        _mat(out, rows, dim, ldo)[...] = v
        if out2:
            _mat(out2, rows, dim, ldo2)[...] = v

    def token_lengths(self, ids, lengths, rows, L, st):
        _flat(lengths, rows, I32)[...] = (_mat(ids, rows, L, L, dtype=I64) != 0).sum(1)

    def feature_lengths(self, x, lengths, rows, L, D, st):
        _flat(lengths, rows, I32)[...] = (_flat(x, rows * L * D).reshape(rows, L, D) != 0).any(2).sum(1)

    def lstm_seq_fwd(self, gx, whh_f, whh_b, bhh_f, bhh_b, lengths, out, hn, sg, sc, R, L, H, dirs, st):
        gxa = _flat(gx, dirs * R * L * 4 * H).reshape(dirs, R, L, 4 * H)
        o = _flat(out, R * L * dirs * H).reshape(R, L, dirs * H)
        o[...] = 0
        lens = np.minimum(_flat(lengths, R, I32), L)
        sga = _flat(sg, dirs * R * L * 4 * H).reshape(dirs, R, L, 4 * H) if sg else None
        sca = _flat(sc, dirs * R * L * H).reshape(dirs, R, L, H) if sc else None
        hna = _flat(hn, dirs * R * H).reshape(dirs, R, H) if hn else None
        for d in range(dirs):
            W = _flat(whh_b if d else whh_f, 4 * H * H).reshape(4 * H, H).astype(np.float64)
            bh = _flat(bhh_b if d else bhh_f, 4 * H).astype(np.float64)
            for r in range(R):
                h, c = np.zeros(H), np.zeros(H)
                for step in range(lens[r]):
                    t = step if d == 0 else lens[r] - 1 - step
                    pre = gxa[d, r, t].astype(np.float64) + bh + W @ h
                    g = np.concatenate([_sig(pre[:2 * H]), np.tanh(pre[2 * H:3 * H]), _sig(pre[3 * H:])])
                    c = g[H:2 * H] * c + g[:H] * g[2 * H:3 * H]
                    h = g[3 * H:] * np.tanh(c)
                    o[r, t, d * H:(d + 1) * H] = h
                    if sga is not None:
                        sga[d, r, t] = g
                    if sca is not None:
                        sca[d, r, t] = c
                if hna is not None:
                    hna[d, r] = h

    def gru_seq_fwd(self, gx, whh_f, whh_b, bhh_f, bhh_b, lengths, out, hn, save, R, L, H, dirs, st):
        gxa = _flat(gx, dirs * R * L * 3 * H).reshape(dirs, R, L, 3 * H)
        o = _flat(out, R * L * dirs * H).reshape(R, L, dirs * H)
        o[...] = 0
        lens = np.minimum(_flat(lengths, R, I32), L)
        sva = _flat(save, dirs * R * L * 4 * H).reshape(dirs, R, L, 4 * H) if save else None
        hna = _flat(hn, dirs * R * H).reshape(dirs, R, H) if hn else None
        for d in range(dirs):
            W = _flat(whh_b if d else whh_f, 3 * H * H).reshape(3 * H, H).astype(np.float64)
            bh = _flat(bhh_b if d else bhh_f, 3 * H).astype(np.float64)
            for r in range(R):
                h = np.zeros(H)
                for step in range(lens[r]):
                    t = step if d == 0 else lens[r] - 1 - step
                    x = gxa[d, r, t].astype(np.float64)
                    gh = W @ h + bh
                    rg, zg = _sig(x[:H] + gh[:H]), _sig(x[H:2 * H] + gh[H:2 * H])
                    ng = np.tanh(x[2 * H:] + rg * gh[2 * H:])
                    h = (1 - zg) * ng + zg * h
                    o[r, t, d * H:(d + 1) * H] = h
                    if sva is not None:
                        sva[d, r, t] = np.concatenate([rg, zg, ng, gh[2 * H:]])
                if hna is not None:
                    hna[d, r] = h

    def gru_seq_bwd(self, dout, save, out, whh_f, whh_b, lengths, dgx, dgh, hprev, R, L, H, dirs, st):
        do = _flat(dout, R * L * dirs * H).reshape(R, L, dirs * H).astype(np.float64)
        sv = _flat(save, dirs * R * L * 4 * H).reshape(dirs, R, L, 4 * H).astype(np.float64)
        o = _flat(out, R * L * dirs * H).reshape(R, L, dirs * H).astype(np.float64)
        dx_ = _flat(dgx, dirs * R * L * 3 * H).reshape(dirs, R, L, 3 * H)
        dh_ = _flat(dgh, dirs * R * L * 3 * H).reshape(dirs, R, L, 3 * H)
        hp_ = _flat(hprev, dirs * R * L * H).reshape(dirs, R, L, H)
        dx_[...] = 0
        dh_[...] = 0
        hp_[...] = 0
        lens = np.minimum(_flat(lengths, R, I32), L)
        for d in range(dirs):
            W = _flat(whh_b if d else whh_f, 3 * H * H).reshape(3 * H, H).astype(np.float64)
            for r in range(R):
                carry = np.zeros(H)
                for step in range(lens[r] - 1, -1, -1):
                    tt = step if d == 0 else lens[r] - 1 - step
                    tp = tt - 1 if d == 0 else tt + 1
                    rg, zg, ng, ghn = (sv[d, r, tt, i * H:(i + 1) * H] for i in range(4))
                    hp = o[r, tp, d * H:(d + 1) * H] if step > 0 else np.zeros(H)
                    hp_[d, r, tt] = hp
                    g = do[r, tt, d * H:(d + 1) * H] + carry
                    dpn = g * (1 - zg) * (1 - ng * ng)
                    dpz = g * (hp - ng) * zg * (1 - zg)
                    dpr = dpn * ghn * rg * (1 - rg)
                    dx_[d, r, tt] = np.concatenate([dpr, dpz, dpn])
                    gh = np.concatenate([dpr, dpz, dpn * rg])
                    dh_[d, r, tt] = gh
                    carry = g * zg + W.T @ gh

    def gru_pointwise_fwd(self, gi, gh, h, ldh, hout, ldo, hout2, ldo2, save, rows, H, st):
        gi_, gh_ = _mat(gi, rows, 3 * H, 3 * H).astype(np.float64), _mat(gh, rows, 3 * H, 3 * H).astype(np.float64)
        hp = _mat(h, rows, H, ldh).astype(np.float64)
        r = _sig(gi_[:, :H] + gh_[:, :H])
        z = _sig(gi_[:, H:2 * H] + gh_[:, H:2 * H])
        ghn = gh_[:, 2 * H:]
        n = np.tanh(gi_[:, 2 * H:] + r * ghn)
        hn = (1 - z) * n + z * hp
        _mat(hout, rows, H, ldo)[...] = hn
        if hout2:
            _mat(hout2, rows, H, ldo2)[...] = hn
        if save:
            _mat(save, rows, 4 * H, 4 * H)[...] = np.concatenate([r, z, n, ghn], 1)

    def gru_step_fwd(self, x, ldx, in_size, h0, ldh0, mask, w_ih, w_hh, b_ih, b_hh, hm, out, ldo, out2, ldo2, save,
                     rows, H, st):
        hp = _mat(h0, rows, H, ldh0).astype(np.float64) * _flat(mask, rows).astype(np.float64)[:, None]
        gi_ = _mat(x, rows, in_size, ldx).astype(np.float64) @ _mat(w_ih, 3 * H, in_size, in_size).astype(np.float64).T \
            + _flat(b_ih, 3 * H).astype(np.float64)
        gh_ = hp @ _mat(w_hh, 3 * H, H, H).astype(np.float64).T + _flat(b_hh, 3 * H).astype(np.float64)
        r = _sig(gi_[:, :H] + gh_[:, :H])
        z = _sig(gi_[:, H:2 * H] + gh_[:, H:2 * H])
        ghn = gh_[:, 2 * H:]
        n = np.tanh(gi_[:, 2 * H:] + r * ghn)
        hn = (1 - z) * n + z * hp
        _mat(out, rows, H, ldo)[...] = hn
        if out2:
            _mat(out2, rows, H, ldo2)[...] = hn
        if hm:
            _mat(hm, rows, H, H)[...] = hp
        if save:
            _mat(save, rows, 4 * H, 4 * H)[...] = np.concatenate([r, z, n, ghn], 1)

    def lstm_pointwise_fwd(self, gi, gh, c, ldc, hout, ldo, cout, ldco, hout2, cout2, ldo2, save, rows, H, st):
        a = _mat(gi, rows, 4 * H, 4 * H).astype(np.float64) + _mat(gh, rows, 4 * H, 4 * H).astype(np.float64)
        i, f, g, o = _sig(a[:, :H]), _sig(a[:, H:2 * H]), np.tanh(a[:, 2 * H:3 * H]), _sig(a[:, 3 * H:])
        cn = f * _mat(c, rows, H, ldc).astype(np.float64) + i * g
        tc = np.tanh(cn)
        _mat(hout, rows, H, ldo)[...] = o * tc
        _mat(cout, rows, H, ldco)[...] = cn
        if hout2:
            _mat(hout2, rows, H, ldo2)[...] = o * tc
            _mat(cout2, rows, H, ldo2)[...] = cn
        if save:
            _mat(save, rows, 5 * H, 5 * H)[...] = np.concatenate([i, f, g, o, tc], 1)

    def transpose_f32(self, src, lds, dst, ldd, rows, cols, round_tf32, st):
        _mat(dst, cols, rows, ldd)[...] = _mat(src, rows, cols, lds).T      # (host model: no TF32 rounding)

    def round_tf32(self, src, dst, n, st):
        _flat(dst, n)[...] = _flat(src, n)

    def mask_rows(self, h, ldh, mask, out, ldo, rows, H, st):
        _mat(out, rows, H, ldo)[...] = _mat(h, rows, H, ldh) * _flat(mask, rows)[:, None]

    def attn_pool_fwd(self, q, ldq, E, lde_r, lde_p, lengths, scale, probs, out, ldo, R, P, D, st):
        e = _ten3(E, R, P, D, lde_r, lde_p).astype(np.float64)
        s = np.einsum("rd,rpd->rp", _mat(q, R, D, ldq).astype(np.float64), e)
        if lengths:
            s = np.where(np.arange(P)[None, :] < _flat(lengths, R, I32)[:, None], s, s - 1e8)
        s = s * scale
        a = np.exp(s - s.max(1, keepdims=True))
        a /= a.sum(1, keepdims=True)
        if probs:
            _mat(probs, R, P, P)[...] = a
        _mat(out, R, D, ldo)[...] = np.einsum("rp,rpd->rd", a, e)

    def mean_positions(self, E, lde_r, lde_p, out, ldo, R, P, D, st):
        _mat(out, R, D, ldo)[...] = _ten3(E, R, P, D, lde_r, lde_p).astype(np.float64).mean(1)

    def spatial_embed_fill(self, table, E, lde_r, lde_p, c_off, R, P, Cs, st):
        t = _flat(table, P * Cs).reshape(Cs, P)  # raw view(1, Cs, 4, 4): [c, s]
        _ten3(E + 4 * c_off, R, P, Cs, lde_r, lde_p)[...] = t.T[None]

    def positions_to_nchw(self, E, lde_r, lde_p, out, ldo, R, P, C, st):
        e = _ten3(E, R, P, C, lde_r, lde_p)
        _mat(out, R, C * P, ldo)[...] = e.transpose(0, 2, 1).reshape(R, C * P)

    def nchw_to_positions(self, x, y, ldy, N, C, P, st):
        xa = _flat(x, N * C * P).reshape(N, C, P)
        _ten3(y, N, P, C, P * ldy, ldy)[...] = xa.transpose(0, 2, 1)

    def log_softmax_rows(self, x, rows, A, st):
        a = _mat(x, rows, A, A)
        v = a.astype(np.float64)
        m = v.max(1, keepdims=True)
        a[...] = v - (m + np.log(np.exp(v - m).sum(1, keepdims=True)))

    def argmax_rows(self, x, out, rows, A, st):
        _flat(out, rows, I64)[...] = _mat(x, rows, A, A).argmax(1)

    # ---- backward --------------------------------------------------------------------
    def relu_bwd(self, d, ldd, y, ldy, rows, cols, st):
        dv = _mat(d, rows, cols, ldd)
        dv[_mat(y, rows, cols, ldy) <= 0] = 0

    def tanh_bwd(self, d, y, n, st):
        dv, yv = _flat(d, n), _flat(y, n)
        dv *= 1 - yv * yv

    def log_softmax_bwd(self, lp, dlp, draw, rows, A, st):
        l, d = _mat(lp, rows, A, A), _mat(dlp, rows, A, A)
        _mat(draw, rows, A, A)[...] = d - np.exp(l) * d.sum(1, keepdims=True)

    def mask_rows_add(self, src, lds, mask, dst, ldd, rows, H, st):
        _mat(dst, rows, H, ldd)[...] += _mat(src, rows, H, lds) * _flat(mask, rows)[:, None]

    def gru_pointwise_bwd(self, dh, lddh, save, hm, dgi, dgh, dhm, rows, H, st):
        s = _mat(save, rows, 4 * H, 4 * H).astype(np.float64)
        r, z, n, ghn = s[:, :H], s[:, H:2 * H], s[:, 2 * H:3 * H], s[:, 3 * H:]
        g = _mat(dh, rows, H, lddh).astype(np.float64)
        hp = _mat(hm, rows, H, H).astype(np.float64)
        dpn = g * (1 - z) * (1 - n * n)
        dpz = g * (hp - n) * z * (1 - z)
        dpr = dpn * ghn * r * (1 - r)
        _mat(dgi, rows, 3 * H, 3 * H)[...] = np.concatenate([dpr, dpz, dpn], 1)
        _mat(dgh, rows, 3 * H, 3 * H)[...] = np.concatenate([dpr, dpz, dpn * r], 1)
        _mat(dhm, rows, H, H)[...] = g * z

    def lstm_pointwise_bwd(self, dh, lddh, dc_in, save, cm, mask, dgates, dc_out, rows, H, st):
        s = _mat(save, rows, 5 * H, 5 * H).astype(np.float64)
        i, f, gg, o, tc = (s[:, k * H:(k + 1) * H] for k in range(5))
        g = _mat(dh, rows, H, lddh).astype(np.float64)
        dct = g * o * (1 - tc * tc)
        if dc_in:
            dct = dct + _mat(dc_in, rows, H, H).astype(np.float64)
        cp = _mat(cm, rows, H, H).astype(np.float64)
        _mat(dgates, rows, 4 * H, 4 * H)[...] = np.concatenate(
            [dct * gg * i * (1 - i), dct * cp * f * (1 - f), dct * i * (1 - gg * gg), g * tc * o * (1 - o)], 1)
        m = _flat(mask, rows)[:, None] if mask else 1.0
        _mat(dc_out, rows, H, H)[...] = dct * f * m

    def attn_pool_bwd(self, q, ldq, E, lde_r, lde_p, probs, scale, dout, lddo, dq, lddq, dE, ldde_r, ldde_p, acc,
                      R, P, D, st):
        e = _ten3(E, R, P, D, lde_r, lde_p).astype(np.float64)
        a = _mat(probs, R, P, P).astype(np.float64)
        do = _mat(dout, R, D, lddo).astype(np.float64)
        qv = _mat(q, R, D, ldq).astype(np.float64)
        da = np.einsum("rpd,rd->rp", e, do)
        dl = a * (da - (a * da).sum(1, keepdims=True)) * scale
        _mat(dq, R, D, lddq)[...] = np.einsum("rp,rpd->rd", dl, e)
        if dE:
            v = a[:, :, None] * do[:, None, :] + dl[:, :, None] * qv[:, None, :]
            de = _ten3(dE, R, P, D, ldde_r, ldde_p)
            de[...] = (de + v) if acc else v

    def lstm_seq_bwd(self, dout, gates, cs, out, whh_f, whh_b, lengths, dgx, hprev, R, L, H, dirs, st):
        do = _flat(dout, R * L * dirs * H).reshape(R, L, dirs * H).astype(np.float64)
        g = _flat(gates, dirs * R * L * 4 * H).reshape(dirs, R, L, 4 * H).astype(np.float64)
        c = _flat(cs, dirs * R * L * H).reshape(dirs, R, L, H).astype(np.float64)
        o = _flat(out, R * L * dirs * H).reshape(R, L, dirs * H).astype(np.float64)
        dg = _flat(dgx, dirs * R * L * 4 * H).reshape(dirs, R, L, 4 * H)
        hp_ = _flat(hprev, dirs * R * L * H).reshape(dirs, R, L, H)
        dg[...] = 0
        hp_[...] = 0
        lens = np.minimum(_flat(lengths, R, I32), L)
        for d in range(dirs):
            W = _flat(whh_b if d else whh_f, 4 * H * H).reshape(4 * H, H).astype(np.float64)
            for r in range(R):
                dh_c, dc_c = np.zeros(H), np.zeros(H)
                for step in range(lens[r] - 1, -1, -1):
                    tt = step if d == 0 else lens[r] - 1 - step
                    tp = tt - 1 if d == 0 else tt + 1
                    has = step > 0
                    ig, fg, gg, og = (g[d, r, tt, i * H:(i + 1) * H] for i in range(4))
                    cp = c[d, r, tp] if has else np.zeros(H)
                    hp = o[r, tp, d * H:(d + 1) * H] if has else np.zeros(H)
                    hp_[d, r, tt] = hp
                    th = np.tanh(c[d, r, tt])
                    dh = do[r, tt, d * H:(d + 1) * H] + dh_c
                    dc = dh * og * (1 - th * th) + dc_c
                    dp = np.concatenate([dc * gg * ig * (1 - ig), dc * cp * fg * (1 - fg), dc * ig * (1 - gg * gg),
                                         dh * th * og * (1 - og)])
                    dc_c = dc * fg
                    dg[d, r, tt] = dp
                    dh_c = W.T @ dp

    def scatter_last_token(self, dh, lddh, lengths, dout, R, L, H, st):
        o = _flat(dout, R * L * H).reshape(R, L, H)
        o[...] = 0
        lens = _flat(lengths, R, I32)
        d = _mat(dh, R, H, lddh)
        for r in range(R):
            if 1 <= lens[r] <= L:
                o[r, lens[r] - 1] = d[r]

    def embedding_bwd(self, ids, d, ldd, dtable, rows, dim, vocab, padding_idx, st):
        i = _flat(ids, rows, I64)
        t = _mat(dtable, vocab, dim, dim)
        ok = (i >= 0) & (i < vocab) & (i != padding_idx)
        np.add.at(t, i[ok], _mat(d, rows, dim, ldd)[ok])

    def prev_action_embed_bwd(self, prev, masks, d1, ld1, d2, ld2, dtable, rows, dim, table_rows, st):
        v = _mat(d1, rows, dim, ld1).copy()
        if d2:
            v = v + _mat(d2, rows, dim, ld2)
        np.add.at(_mat(dtable, table_rows, dim, dim), self._pa_index(prev, masks, rows, table_rows), v)

    def spatial_embed_grad(self, dE, lde_r, lde_p, c_off, dflat, ldf, dmean, ldm, dtable, accumulate, R, P, Cs, st):
        acc = np.zeros((Cs, P))
        if dE:
            acc += _ten3(dE + 4 * c_off, R, P, Cs, lde_r, lde_p).astype(np.float64).sum(0).T
        if dflat:
            acc += _mat(dflat, R, Cs * P, ldf).astype(np.float64).sum(0).reshape(Cs, P)
        if dmean:
            acc += (_mat(dmean, R, Cs, ldm).astype(np.float64).sum(0) / P)[:, None]
        t = _flat(dtable, Cs * P).reshape(Cs, P)
        t[...] = (t + acc) if accumulate else acc

    def il_loss(self, logp, actions, weights, hat, progress, mode, alpha, inv_accum, T, N, A, d_raw, d_hat, losses, st):
        R = T * N
        lp = _mat(logp, R, A, A).astype(np.float64)
        a = _flat(actions, R, I64)
        w = _flat(weights, R).astype(np.float64)
        wsum = w.reshape(T, N).sum(0)
        n = np.arange(R) % N
        ce = -lp[np.arange(R), a]
        action_loss = (w * ce / wsum[n]).sum() / N
        valid = (w > 0).astype(np.float64)
        M = valid.sum()
        coef = w / (wsum[n] * N) * inv_accum
        onehot = np.zeros((R, A))
        onehot[np.arange(R), a] = 1
        _mat(d_raw, R, A, A)[...] = coef[:, None] * (np.exp(lp) - onehot)
        aux = 0.0
        if mode != 0:
            h, p = _flat(hat, R).astype(np.float64), _flat(progress, R).astype(np.float64)
            if mode == 1:
                l, g = (h - p) ** 2, 2 * (h - p)
            else:
                l, g = h * h - 2 * h * p.mean() + (p * p).mean(), 2 * (h - p.mean())
            aux = alpha * (valid * l).sum() / M
            _flat(d_hat, R)[...] = valid * alpha * g / M * inv_accum
        _flat(losses, 3)[...] = [(action_loss + aux) * inv_accum, action_loss, aux]

    def adam_step(self, p, g, m, v, n, lr, b1, b2, eps, step, step_dev, grad_scale, st):
        if step == 0:
            step = int(_flat(step_dev, 1, I32)[0])
        pv, gv, mv, vv = (_flat(x, n) for x in (p, g, m, v))
        gi = gv.astype(np.float64) * grad_scale
        mi = b1 * mv + (1 - b1) * gi
        vi = b2 * vv + (1 - b2) * gi * gi
        mv[...], vv[...] = mi, vi
        bc1, bc2 = 1 - b1 ** step, math.sqrt(1 - b2 ** step)
        pv[...] = pv - (lr / bc1) * (mi / (np.sqrt(vi) / bc2 + eps))

    def counter_add(self, counter, value, st):
        _flat(counter, 1, I32)[0] += value

    def axpy(self, y, x, alpha, n, st):
        _flat(y, n)[...] += alpha * _flat(x, n)


# ======================================================================================
# waypoint path (csrc/waypoint.cu)
# ======================================================================================
_S2PI = math.sqrt(2 * math.pi)


def _phi(x):
    return np.exp(-0.5 * x * x) / _S2PI


def _Phi(x):
    from scipy.special import erf

    return 0.5 * (1 + erf(x / math.sqrt(2)))


def _tn_stats(mu, sg, lo, hi):
    a, b = (lo - mu) / sg, (hi - mu) / sg
    pa, pb = _phi(a), _phi(b)
    Z = _Phi(b) - _Phi(a)
    dz = (pb - pa) / Z
    return dict(a=a, b=b, pa=pa, pb=pb, Z=Z, mean=mu - sg * dz,
                var=sg * sg * (1 - (b * pb - a * pa) / Z - dz * dz),
                ent=0.5 * math.log(2 * math.pi * math.e) + np.log(sg * Z) + (a * pa - b * pb) / (2 * Z))


def _tn_logp(y, mu, sg, Z):
    u = (y - mu) / sg
    return -0.5 * u * u - np.log(sg * _S2PI * Z)


def _tn_grads(t, y, mu, sg):
    inv = 1 / sg
    u = (y - mu) * inv
    dZm = (t["pa"] - t["pb"]) * inv
    N = t["a"] * t["pa"] - t["b"] * t["pb"]
    dZs = N * inv
    dlp_dmu = u * inv - dZm / t["Z"]
    dlp_ds = u * u * inv - inv - dZs / t["Z"]
    dNm = ((1 - t["b"] ** 2) * t["pb"] - (1 - t["a"] ** 2) * t["pa"]) * inv
    dNs = (t["b"] * (1 - t["b"] ** 2) * t["pb"] - t["a"] * (1 - t["a"] ** 2) * t["pa"]) * inv
    dH_dmu = dZm / t["Z"] + dNm / (2 * t["Z"]) - N * dZm / (2 * t["Z"] ** 2)
    dH_ds = inv + dZs / t["Z"] + dNs / (2 * t["Z"]) - N * dZs / (2 * t["Z"] ** 2)
    return dlp_dmu, dlp_ds * 0.5 * inv, dH_dmu, dH_ds * 0.5 * inv


def _host_floats(ptr, n):
    return [float(v) for v in _flat(ptr, n)]


def _act_consts(ptr):
    v = _host_floats(ptr, 11)
    return dict(d_lo=v[0], d_hi=v[1], o_lim=v[2], pd=v[3] != 0, po=v[4] != 0, nd=int(v[5]), no=int(v[6]), dist_a=v[7],
                dist_b=v[8], off_a=v[9], off_b=v[10])


class _Comp:
    """One action component at the selected pano of each row: TruncatedNormal(v1[idx], sqrt(v2[idx])) or a
    Categorical over the ``ncls`` logits v1[idx, :] (mirrors ``Comp`` in csrc/waypoint.cu)."""

    def __init__(self, ncls, v1, v2, R, P, rows, idx, lo, hi):
        self.ncls, self.R, self.P, self.rows, self.idx, self.lo, self.hi = ncls, R, P, rows, idx, lo, hi
        if ncls:
            lg = _flat(v1, R * P * ncls).reshape(R, P, ncls).astype(np.float64)[rows, idx]
            self.l = lg - (lg.max(1, keepdims=True) + np.log(np.exp(lg - lg.max(1, keepdims=True)).sum(1, keepdims=True)))
        else:
            self.mu = _mat(v1, R, P, P).astype(np.float64)[rows, idx]
            self.sg = np.sqrt(_mat(v2, R, P, P).astype(np.float64)[rows, idx])
            self.t = _tn_stats(self.mu, self.sg, lo, hi)

    def _k(self, y):
        return np.clip(np.asarray(y).astype(np.int64), 0, self.ncls - 1)

    def log_prob(self, y):
        if self.ncls:
            return self.l[np.arange(len(self.l)), self._k(y)]
        return _tn_logp(y, self.mu, self.sg, self.t["Z"])

    def ent(self):
        return -(np.exp(self.l) * self.l).sum(1) if self.ncls else self.t["ent"]

    def var(self):
        return np.full(len(self.l), np.nan) if self.ncls else self.t["var"]

    def mode(self):
        return np.exp(self.l).argmax(1).astype(np.float64) if self.ncls else self.mu

    def sample(self, rng):
        if self.ncls:
            return float(min(np.searchsorted(np.cumsum(np.exp(self.l[0])), rng.random(), side="right"), self.ncls - 1))
        while True:
            s = rng.normal(self.mu[0], self.sg[0])
            if self.lo <= s <= self.hi:
                return s

    def grads(self, y, mask, mask_h, g_lp, g_h, o1, o2):
        R, P, w = self.R, self.P, max(self.ncls, 1)
        g1 = _flat(o1, R * P * w).reshape(R, P, w)
        g1[...] = 0
        if self.ncls:
            p = np.exp(self.l)
            onehot = np.zeros_like(p)
            onehot[np.arange(R), self._k(y)] = 1
            H = self.ent()
            mh = np.broadcast_to(np.asarray(mask_h, dtype=np.float64), (R,))
            g1[self.rows, self.idx] = (mask[:, None] * np.asarray(g_lp)[:, None] * (onehot - p)
                                       + mh[:, None] * np.asarray(g_h).reshape(-1, 1) * (-p * (self.l + H[:, None])))
            return
        g2 = _mat(o2, R, P, P)
        g2[...] = 0
        dlm, dlv, dhm, dhv = _tn_grads(self.t, y, self.mu, self.sg)
        g1[self.rows, self.idx, 0] = mask * g_lp * dlm + mask_h * g_h * dhm
        g2[self.rows, self.idx] = mask * g_lp * dlv + mask_h * g_h * dhv


class WaypointOps:
    def wp_stack_frames(self, cur, hist, masks, out, B, P, frame_bytes, st):
        o = _flat(out, B * (P + 1) * frame_bytes, np.uint8)
        o[: B * P * frame_bytes] = _flat(cur, B * P * frame_bytes, np.uint8)
        h = _flat(hist, B * frame_bytes, np.uint8).reshape(B, frame_bytes)
        m = _flat(masks, B) != 0
        o[B * P * frame_bytes:].reshape(B, frame_bytes)[...] = np.where(m[:, None], h, 0)

    def broadcast_positions(self, E, lde_r, lde_p, R, P, C, st):
        e = _ten3(E, R, P, C, lde_r, lde_p)
        e[:, 1:, :] = e[:, :1, :]

    def wp_prev_action(self, pano, offset, distance, masks, out, ldo, out2, ldo2, rows, panos, off_a, off_b, dist_a,
                       dist_b, st):
        h = _flat(pano, rows, I64).astype(np.float64) * (2 * math.pi / panos)
        m = _flat(masks, rows).astype(np.float64)
        v = np.stack([np.sin(h), np.cos(h), off_a + off_b * _flat(offset, rows), dist_a + dist_b * _flat(distance, rows)],
                     1) * m[:, None]
        _mat(out, rows, 4, ldo)[...] = v
        if out2:
            _mat(out2, rows, 4, ldo2)[...] = v

    def wp_inst_attn_fwd(self, q, ldq, bk, instr, lengths, scale, out, ldo, save, R, L, D, Dk, st):
        lens = np.minimum(_flat(lengths, R, I32), L)
        lmax = int(lens.max())
        c = (_mat(q, R, Dk, ldq).astype(np.float64) @ _flat(bk, Dk).astype(np.float64)) * scale
        x = _flat(instr, R * L * D).reshape(R, L, D).astype(np.float64)
        sv = _mat(save, R, 2, 2)
        for r in range(R):
            npad = lmax - lens[r]
            z = lens[r] + npad * math.exp(c[r])
            a_v = 1.0 / z
            _mat(out, R, D, ldo)[r] = a_v * x[r, : lens[r]].sum(0)
            sv[r] = [a_v, npad * math.exp(c[r]) / z]

    def wp_inst_attn_bwd(self, save, out, ldo, dout, lddo, lengths, scale, dcs, d_instr, R, L, D, st):
        sv = _mat(save, R, 2, 2).astype(np.float64)
        o, do = _mat(out, R, D, ldo).astype(np.float64), _mat(dout, R, D, lddo).astype(np.float64)
        _flat(dcs, R)[...] = -(o * do).sum(1) * sv[:, 1] * scale
        lens = np.minimum(_flat(lengths, R, I32), L)
        di = _flat(d_instr, R * L * D).reshape(R, L, D)
        di[...] = 0
        for r in range(R):
            di[r, : lens[r]] = sv[r, 0] * do[r]

    def wp_scramble_query(self, q, out, n_in, P, st):
        _flat(out, n_in * P)[...] = np.repeat(_flat(q, n_in), P)

    def wp_scramble_query_bwd(self, dout, dq, n_in, P, st):
        _flat(dq, n_in)[...] = _flat(dout, n_in * P).reshape(n_in, P).astype(np.float64).sum(1)

    def layernorm_fwd(self, x, ldx, gamma, beta, y, ldy, save, R, D, eps, st):
        xv = _mat(x, R, D, ldx).astype(np.float64)
        mean = xv.mean(1, keepdims=True)
        rstd = 1 / np.sqrt(((xv - mean) ** 2).mean(1, keepdims=True) + eps)
        _mat(y, R, D, ldy)[...] = (xv - mean) * rstd * _flat(gamma, D) + _flat(beta, D)
        if save:
            _mat(save, R, 2, 2)[...] = np.concatenate([mean, rstd], 1)

    def layernorm_bwd(self, dy, lddy, x, ldx, gamma, save, dx, lddx, prod, ldp, R, D, st):
        sv = _mat(save, R, 2, 2).astype(np.float64)
        xh = (_mat(x, R, D, ldx).astype(np.float64) - sv[:, :1]) * sv[:, 1:]
        d = _mat(dy, R, D, lddy).astype(np.float64)
        g = d * _flat(gamma, D)
        _mat(dx, R, D, lddx)[...] = sv[:, 1:] * (g - g.mean(1, keepdims=True) - xh * (g * xh).mean(1, keepdims=True))
        _mat(prod, R, D, ldp)[...] = d * xh

    def wp_dot_frames_fwd(self, F, ldf_r, ldf_p, xs, ldx, out, ldo, B, P, D, st):
        f = _ten3(F, B, P, D, ldf_r, ldf_p).astype(np.float64)
        _mat(out, B, P, ldo)[...] = np.einsum("bpd,bd->bp", f, _mat(xs, B, D, ldx).astype(np.float64))

    def wp_dot_frames_bwd(self, F, ldf_r, ldf_p, xs, ldx, dout, lddo, dF, lddf_r, lddf_p, acc, dxs, lddx, B, P, D, st):
        f = _ten3(F, B, P, D, ldf_r, ldf_p).astype(np.float64)
        g = _mat(dout, B, P, lddo).astype(np.float64)
        x = _mat(xs, B, D, ldx).astype(np.float64)
        v = g[:, :, None] * x[:, None, :]
        df = _ten3(dF, B, P, D, lddf_r, lddf_p)
        df[...] = (df + v) if acc else v
        _mat(dxs, B, D, lddx)[...] = np.einsum("bp,bpd->bd", g, f)

    def broadcast_rows(self, x, ldx, out, ldo_r, ldo_p, B, P, D, st):
        _ten3(out, B, P, D, ldo_r, ldo_p)[...] = _mat(x, B, D, ldx)[:, None, :]

    def sum_positions(self, E, lde_r, lde_p, out, ldo, acc, R, P, D, st):
        s = _ten3(E, R, P, D, lde_r, lde_p).astype(np.float64).sum(1)
        o = _mat(out, R, D, ldo)
        o[...] = (o + s) if acc else s

    def wp_head_act_fwd(self, raw, d1, d2, o1, o2, n, consts, st):
        c = _host_floats(consts, 8)
        r = _mat(raw, n, 4, 4).astype(np.float64)
        _flat(d1, n)[...] = c[0] * _sig(r[:, 0]) + c[1]
        _flat(d2, n)[...] = c[2] * _sig(r[:, 1]) + c[3]
        _flat(o1, n)[...] = c[4] * np.tanh(r[:, 2] * c[5])
        _flat(o2, n)[...] = c[6] * _sig(r[:, 3]) + c[7]

    def wp_head_act_bwd(self, raw, g_d1, g_d2, g_o1, g_o2, g_raw, n, consts, st):
        c = _host_floats(consts, 8)
        r = _mat(raw, n, 4, 4).astype(np.float64)
        s0, s1, s3, t = _sig(r[:, 0]), _sig(r[:, 1]), _sig(r[:, 3]), np.tanh(r[:, 2] * c[5])
        _mat(g_raw, n, 4, 4)[...] = np.stack([
            _flat(g_d1, n) * c[0] * s0 * (1 - s0), _flat(g_d2, n) * c[2] * s1 * (1 - s1),
            _flat(g_o1, n) * c[4] * (1 - t * t) * c[5], _flat(g_o2, n) * c[6] * s3 * (1 - s3)], 1)

    def wp_act(self, logp, o1, o2, d1, d2, pano_stop, fout, B, P, deterministic, seed, offset, consts, st):
        c = _act_consts(consts)
        lp = _mat(logp, B, P + 1, P + 1).astype(np.float64)
        rng = np.random.default_rng(seed + offset)
        f = _mat(fout, B, 10, 10)
        for b in range(B):
            if deterministic:
                ps = int(np.exp(lp[b]).argmax())
            else:
                ps = int(min(np.searchsorted(np.cumsum(np.exp(lp[b])), rng.random(), side="right"), P))
            pano, go = ps % P, float(ps != P)
            idx = np.array([pano])
            cd = _Comp(c["nd"], d1, d2, B, P, np.array([b]), idx, c["d_lo"], c["d_hi"])
            co = _Comp(c["no"], o1, o2, B, P, np.array([b]), idx, -c["o_lim"], c["o_lim"])
            dist = cd.mode()[0] if deterministic else cd.sample(rng)
            ofs = co.mode()[0] if deterministic else co.sample(rng)
            lp_d, lp_o = cd.log_prob(np.array([dist]))[0], co.log_prob(np.array([ofs]))[0]
            a_d, a_o = c["dist_a"] + c["dist_b"] * dist, c["off_a"] + c["off_b"] * ofs
            var_d, var_o = cd.var()[0], co.var()[0]
            if not c["po"]:
                a_o = ofs = var_o = 0.0
            if not c["pd"]:
                a_d, dist, var_d = 0.25, (0.0 if c["nd"] else 0.25), 0.0
            theta = (pano * (2 * math.pi / P) + a_o) % (2 * math.pi)
            alp = lp[b, ps] + (go * lp_d if c["pd"] else 0.0) + (go * lp_o if c["po"] else 0.0)
            _flat(pano_stop, B, I64)[b] = ps
            f[b] = [ofs, dist, alp, var_d, var_o, co.mode()[0], cd.mode()[0], a_d if go else 0.0, theta if go else 0.0,
                    1.0 - go]

    def wp_ppo_loss(self, logp, o1, o2, d1, d2, values, a_pano, a_offset, a_distance, v_old, ret, lp_old, adv, R, P,
                    act_consts, ppo_consts, eval_out, losses, d_raw, d_o1, d_o2, d_d1, d_d2, d_values, st):
        c = _act_consts(act_consts)
        lp = _mat(logp, R, P + 1, P + 1).astype(np.float64)
        a = _flat(a_pano, R).astype(np.int64)
        idx, go = a % P, (a != P).astype(np.float64)
        dm, om = go * float(c["pd"]), go * float(c["po"])
        rows = np.arange(R)
        p = np.exp(lp)
        Hp = -(p * lp).sum(1)
        cd = _Comp(c["nd"], d1, d2, R, P, rows, idx, c["d_lo"], c["d_hi"])
        co = _Comp(c["no"], o1, o2, R, P, rows, idx, -c["o_lim"], c["o_lim"])
        y_d, y_o = _flat(a_distance, R).astype(np.float64), _flat(a_offset, R).astype(np.float64)
        alp = lp[rows, a] + dm * cd.log_prob(y_d) + om * co.log_prob(y_o)
        # W4: a categorical head's entropy is weighted by mean(mask) (the reference's (R,1)*(R,) broadcast)
        dmh = go.mean() * float(c["pd"]) if c["nd"] else dm
        omh = go.mean() * float(c["po"]) if c["no"] else om
        Ho, Hd = omh * co.ent(), dmh * cd.ent()
        if eval_out:
            _mat(eval_out, R, 4, 4)[...] = np.stack([alp, Hp, co.ent() if c["no"] else Ho, cd.ent() if c["nd"] else Hd], 1)
        if not losses:
            return
        clip, vcoef, ecoef, oreg, cp, cof, cdi, clipped = _host_floats(ppo_consts, 8)
        ratio = np.exp(alp - _flat(lp_old, R))
        A = _flat(adv, R).astype(np.float64)
        s1, s2 = ratio * A, np.clip(ratio, 1 - clip, 1 + clip) * A
        v, vo, rt = (_flat(x, R).astype(np.float64) for x in (values, v_old, ret))
        diff = v - vo
        vc = vo + np.clip(diff, -clip, clip)
        l1, l2 = (v - rt) ** 2, (vc - rt) ** 2
        value_loss = (0.5 * np.maximum(l1, l2) if clipped else 0.5 * l1).mean() * vcoef
        action_loss = -np.minimum(s1, s2).mean()
        entropy_loss = (cp * Hp + cof * Ho + cdi * Hd).mean() * ecoef
        offset_loss = oreg * np.abs(c["off_a"] + c["off_b"] * y_o).mean()
        _flat(losses, 8)[...] = [value_loss + action_loss + offset_loss - entropy_loss, value_loss, action_loss,
                                 entropy_loss, offset_loss, Hp.mean(), Ho.mean(), Hd.mean()]
        if not d_raw:
            return
        inside = (ratio >= 1 - clip) & (ratio <= 1 + clip)
        g_lp = np.where(inside | (s1 < s2), -ratio * A / R, 0.0)
        g_ent = -ecoef / R
        onehot = np.zeros((R, P + 1))
        onehot[rows, a] = 1
        _mat(d_raw, R, P + 1, P + 1)[...] = g_lp[:, None] * (onehot - p) + g_ent * cp * (-p * (lp + Hp[:, None]))
        cd.grads(y_d, dm, dmh, g_lp, g_ent * cdi, d_d1, d_d2)
        co.grads(y_o, om, omh, g_lp, g_ent * cof, d_o1, d_o2)
        dv = np.where((l1 >= l2) | (not clipped), v - rt, np.where(np.abs(diff) < clip, vc - rt, 0.0))
        _flat(d_values, R)[...] = vcoef * dv / R

    def clip_grad_norm(self, g, n, max_norm, pre_scale, ws, ws_bytes, st):
        gv = _flat(g, n)
        norm = math.sqrt(float((gv.astype(np.float64) ** 2).sum())) * pre_scale
        f = min(1.0, max_norm / (norm + 1e-6)) * pre_scale
        w = _flat(ws, 1026)
        w[1024], w[1025] = norm, f
        gv *= np.float32(f)

    def gae_returns(self, rewards, value_preds, masks, next_value, returns, advantages, T, N, gamma, tau, use_gae, st):
        rw = _flat(rewards, T * N).reshape(T, N)
        vp = _flat(value_preds, (T + 1) * N).reshape(T + 1, N)
        m = _flat(masks, (T + 1) * N, np.uint8).reshape(T + 1, N).astype(F32)
        rt = _flat(returns, (T + 1) * N).reshape(T + 1, N)
        nv = _flat(next_value, N)
        if use_gae:
            vp[T] = nv
            gae = np.zeros(N, F32)
            for t in range(T - 1, -1, -1):
                delta = rw[t] + np.float32(gamma) * vp[t + 1] * m[t + 1] - vp[t]
                gae = delta + np.float32(gamma * tau) * m[t + 1] * gae
                rt[t] = gae + vp[t]
        else:
            rt[T] = nv
            for t in range(T - 1, -1, -1):
                rt[t] = rt[t + 1] * np.float32(gamma) * m[t + 1] + rw[t]
        if advantages:
            _flat(advantages, T * N).reshape(T, N)[...] = rt[:T] - vp[:T]


for _name in [n for n in vars(WaypointOps) if not n.startswith("_")]:
    setattr(Emulator, _name, getattr(WaypointOps, _name))


def _wp_eval_bwd(self, logp, o1, o2, d1, d2, a_pano, a_offset, a_distance, upstream, R, P, act_consts, d_raw, d_o1,
                 d_o2, d_d1, d_d2, st):
    c = _act_consts(act_consts)
    lp = _mat(logp, R, P + 1, P + 1).astype(np.float64)
    a = _flat(a_pano, R).astype(np.int64)
    idx, go = a % P, (a != P).astype(np.float64)
    dm, om = go * float(c["pd"]), go * float(c["po"])
    rows = np.arange(R)
    p = np.exp(lp)
    Hp = -(p * lp).sum(1)
    up = _mat(upstream, R, 4, 4).astype(np.float64)
    onehot = np.zeros((R, P + 1))
    onehot[rows, a] = 1
    _mat(d_raw, R, P + 1, P + 1)[...] = up[:, 0:1] * (onehot - p) + up[:, 1:2] * (-p * (lp + Hp[:, None]))
    cd = _Comp(c["nd"], d1, d2, R, P, rows, idx, c["d_lo"], c["d_hi"])
    co = _Comp(c["no"], o1, o2, R, P, rows, idx, -c["o_lim"], c["o_lim"])
    cd.grads(_flat(a_distance, R).astype(np.float64), dm, 1.0 if c["nd"] else dm, up[:, 0], up[:, 3], d_d1, d_d2)
    co.grads(_flat(a_offset, R).astype(np.float64), om, 1.0 if c["no"] else om, up[:, 0], up[:, 2], d_o1, d_o2)


Emulator.wp_eval_bwd = _wp_eval_bwd


def _inflection_weights(self, actions, lengths, weights, T, N, coef, st):
    a = _flat(actions, T * N, I64).reshape(T, N)
    w = _flat(weights, T * N).reshape(T, N)
    infl = np.ones((T, N), bool)
    infl[1:] = a[1:] != a[:-1]
    v = np.where(infl, coef, 1.0)
    if lengths:
        v = v * (np.arange(T)[:, None] < _flat(lengths, N, I32)[None, :])
    w[...] = v


Emulator.inflection_weights = _inflection_weights


def _resize_area_crop(self, x, y, dtype, y_frame_stride, N, H, W, C, RH, RW, y0, x0, CH, CW, st):
    dt = {0: F32, 2: np.uint8}[dtype]
    xin = _flat(x, N * H * W * C, dt).reshape(N, H, W, C).astype(np.float32)
    stride = y_frame_stride or CH * CW * C
    out = _flat(y, (N - 1) * stride + CH * CW * C, dt)
    for oy in range(CH):
        ry = oy + y0
        h0, h1 = (ry * H) // RH, ((ry + 1) * H + RH - 1) // RH
        for ox in range(CW):
            rx = ox + x0
            w0, w1 = (rx * W) // RW, ((rx + 1) * W + RW - 1) // RW
            m = xin[:, h0:h1, w0:w1, :].sum((1, 2), dtype=np.float32) / np.float32((h1 - h0) * (w1 - w0))
            for n in range(N):
                o = n * stride + (oy * CW + ox) * C
                out[o:o + C] = m[n].astype(dt)


Emulator.resize_area_crop = _resize_area_crop


# ======================================================================================
# conv trunks, fp32 precision mode only (``net.precision = "fp32"``: vlnce_conv2d_ref + the normalisation /
# pooling kernels).  Modelled with torch CPU ops on tensors that alias the caller's buffers, so the trunk launch
# plans (block wiring, residual / downsample routing, folded BatchNorm, GroupNorm, stem scaling) can be executed on
# the host as well.  The 16-bit tensor-core paths (operand packing, TMA tiles) are GPU-only.
# ======================================================================================
def _tt(ptr, shape, code=0):
    import torch

    npdt = {0: F32, 2: np.uint8}[code]
    n = int(np.prod(shape))
    return torch.from_numpy(_flat(ptr, n, npdt)).view(*shape)


class TrunkOps:
    def conv2d_ref(self, x, x_dtype, w, w_dtype, scale, shift, residual, res_dtype, y, y_dtype, N, H, W, C, K, R, S,
                   stride, pad, relu, st):
        import torch
        import torch.nn.functional as Fn

        assert w_dtype == 0 and y_dtype == 0, "fp32 mode only"
        xi = _tt(x, (N, H, W, C), x_dtype).float().permute(0, 3, 1, 2)
        wt = _tt(w, (K, R, S, C)).permute(0, 3, 1, 2)
        out = Fn.conv2d(xi.contiguous(), wt.contiguous(), stride=stride, padding=pad)   # fp32 like the kernel
        if scale:
            out = out * _tt(scale, (K,)).view(1, K, 1, 1) + _tt(shift, (K,)).view(1, K, 1, 1)
        P, Q = out.shape[2], out.shape[3]
        if residual:
            out = out + _tt(residual, (N, P, Q, K), res_dtype).float().permute(0, 3, 1, 2)
        if relu:
            out = out.clamp_min(0)
        _tt(y, (N, P, Q, K)).copy_(out.permute(0, 2, 3, 1).float())

    def conv2d_warp(self, x, w, scale, shift, residual, y, N, H, W, C, K, R, S, stride, pad, relu, st):
        self.conv2d_ref(x, 0, w, 0, scale, shift, residual, 0, y, 0, N, H, W, C, K, R, S, stride, pad, relu, st)

    def avgpool2x2(self, x, y, N, H, W, C, st):
        import torch.nn.functional as Fn

        _tt(y, (N, H // 2, W // 2, C)).copy_(Fn.avg_pool2d(_tt(x, (N, H, W, C)).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))

    def maxpool3x3s2(self, x, y, dtype, N, H, W, C, st):
        import torch.nn.functional as Fn

        P, Q = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        out = Fn.max_pool2d(_tt(x, (N, H, W, C), dtype).permute(0, 3, 1, 2), 3, 2, 1)
        _tt(y, (N, P, Q, C), dtype).copy_(out.permute(0, 2, 3, 1))

    def adaptive_avgpool(self, x, dtype, y, ldy, N, H, W, C, O, st):
        import torch.nn.functional as Fn

        out = Fn.adaptive_avg_pool2d(_tt(x, (N, H, W, C), dtype).permute(0, 3, 1, 2).double(), (O, O))   # (N,C,O,O)
        _ten3(y, N, O * O, C, O * O * ldy, ldy)[...] = out.flatten(2).permute(0, 2, 1).float().numpy()

    def groupnorm_fused(self, x, gamma, beta, residual, y, y_dtype, ws, wsb, sc, sh, N, HW, C, groups, eps, relu, st):
        import torch
        import torch.nn.functional as Fn

        assert y_dtype == 0
        xi = _tt(x, (N, HW, C)).double().permute(0, 2, 1)     # (N,C,HW)
        out = Fn.group_norm(xi, groups, _tt(gamma, (C,)).double(), _tt(beta, (C,)).double(), eps)
        if residual:
            out = out + _tt(residual, (N, HW, C)).double().permute(0, 2, 1)
        if relu & 1:   # bit 1 = round the output to TF32 (a no-op for the fp32 host model)
            out = out.clamp_min(0)
        _tt(y, (N, HW, C)).copy_(out.permute(0, 2, 1).float())

    def batchnorm_train_stats(self, x, dtype, gamma, beta, rmean, rvar, momentum, scale, shift, ws, rows, C, eps, st):
        import torch

        xi = _tt(x, (rows, C), dtype).double()
        mean, var = xi.mean(0), xi.var(0, unbiased=False)
        g, b = _tt(gamma, (C,)).double(), _tt(beta, (C,)).double()
        s = g / torch.sqrt(var + eps)
        _tt(scale, (C,)).copy_(s.float())
        _tt(shift, (C,)).copy_((b - mean * s).float())
        rm, rv = _tt(rmean, (C,)), _tt(rvar, (C,))
        rm.copy_(((1 - momentum) * rm.double() + momentum * mean).float())
        rv.copy_(((1 - momentum) * rv.double() + momentum * xi.var(0, unbiased=True)).float())

    def affine_apply(self, x, x_dtype, scale, shift, scale_stride_n, residual, y, y_dtype, N, HW, C, relu, st):
        assert x_dtype == 0 and y_dtype == 0
        xi = _tt(x, (N, HW, C)).double()
        if scale_stride_n:
            sc, sh = _tt(scale, (N, 1, C)).double(), _tt(shift, (N, 1, C)).double()
        else:
            sc, sh = _tt(scale, (1, 1, C)).double(), _tt(shift, (1, 1, C)).double()
        out = xi * sc + sh
        if residual:
            out = out + _tt(residual, (N, HW, C)).double()
        if relu:
            out = out.clamp_min(0)
        _tt(y, (N, HW, C)).copy_(out.float())


for _name in [n for n in vars(TrunkOps) if not n.startswith("_")]:
    setattr(Emulator, _name, getattr(TrunkOps, _name))


# ======================================================================================
# 16-bit tensor-core paths (the default ``precision="bf16"``): functional models of the tcgen05 / TMA conv entry
# points and the 16-bit variants of the normalisation / pooling kernels.  Operands are converted to fp32, the
# convolution runs in fp32 (the kernels accumulate in fp32 in TMEM) and the result is rounded to the output dtype.
# What this checks on the host is the operand PACKING and routing done by trunk.py: KRSC weights, the 2x2
# space-to-depth stem (cell image + re-indexed 4x4 weights), the padded im2col depth stem, gemm_c calls, residual
# and norm wiring, fp16-vs-bf16 trunk dtypes.
# ======================================================================================
def _tt16(ptr, shape, code):
    """torch tensor aliasing caller memory for dtype codes F32=0, BF16=1, U8=2, F16=4."""
    import torch

    n = int(np.prod(shape))
    if code in (0, 2):
        return _tt(ptr, shape, code)
    ct = ctypes.c_uint16
    raw = np.ctypeslib.as_array(ctypes.cast(int(ptr), ctypes.POINTER(ct)), shape=(n,))
    t = torch.from_numpy(raw.view(np.int16))
    return t.view(torch.bfloat16 if code == 1 else torch.float16).view(*shape)


def _epilogue(out, scale, shift, residual, res_code, relu, y, y_code, shape):
    """out (N,K,P,Q) fp32 -> y (N,P,Q,K)"""
    N, K, P, Q = out.shape
    if scale:
        out = out * _tt(scale, (K,)).view(1, K, 1, 1) + _tt(shift, (K,)).view(1, K, 1, 1)
    if residual:
        out = out + _tt16(residual, (N, P, Q, K), res_code).float().permute(0, 3, 1, 2)
    if relu & 1:   # (bit 1 = keep the plain fp32 result: GEMM use of the TF32 tiles)
        out = out.clamp_min(0)
    yt = _tt16(y, (N, P, Q, K), y_code)
    yt.copy_(out.permute(0, 2, 3, 1).to(yt.dtype))


class TensorCoreOps:
    def conv2d_bf16(self, x, w, ab, scale, shift, residual, y, y_dtype, N, H, W, C, K, R, S, stride, pad, relu, code, st):
        import torch.nn.functional as Fn

        xi = _tt16(x, (N, H, W, C), ab).float().permute(0, 3, 1, 2).contiguous()
        wt = _tt16(w, (K, R, S, C), ab).float().permute(0, 3, 1, 2).contiguous()
        _epilogue(Fn.conv2d(xi, wt, stride=stride, padding=pad), scale, shift, residual, ab, relu, y, y_dtype, None)

    def conv_tma(self, x, w, ab, scale, shift, residual, y, y_dtype, N, H, W, C, K, R, relu, code, st):
        self.conv2d_bf16(x, w, ab, scale, shift, residual, y, y_dtype, N, H, W, C, K, R, R, 1, R // 2, relu, code, st)

    def conv_s2_tma(self, x, w, ab, scale, shift, y, y_dtype, N, H, W, C, K, R, relu, code, st):
        self.conv2d_bf16(x, w, ab, scale, shift, 0, y, y_dtype, N, H, W, C, K, R, R, 2, R // 2, relu, code, st)

    def conv3x3_halo(self, x, w, ab, scale, shift, residual, y, y_dtype, N, H, W, C, K, relu, code, st):
        self.conv2d_bf16(x, w, ab, scale, shift, residual, y, y_dtype, N, H, W, C, K, 3, 3, 1, 1, relu, code, st)

    def stem_s2d(self, x, x_code, cells, c_code, N, H, W, st):
        import torch

        P, Q = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        Ys, Xs = P + 3, Q + 3
        img = _tt16(x, (N, H, W, 3), x_code).float()
        pad = torch.zeros(N, 2 * Ys, 2 * Xs, 3)
        # cell (Y,X) holds pixels (2Y+dy-3, 2X+dx-3): shift the image by +3 into a zero canvas
        hh, ww = min(H, 2 * Ys - 3), min(W, 2 * Xs - 3)
        pad[:, 3:3 + hh, 3:3 + ww] = img[:, :hh, :ww]
        c = pad.view(N, Ys, 2, Xs, 2, 3).permute(0, 1, 3, 2, 4, 5).reshape(N, Ys, Xs, 12)
        out = _tt16(cells, (N, Ys, Xs, 16), c_code)
        out.zero_()
        out[..., :12] = c.to(out.dtype)

    def conv_stem_tma(self, cells, w, w_code, scale, shift, y, y_dtype, N, P, Q, K, relu, code, st):
        import torch.nn.functional as Fn

        ci = _tt16(cells, (N, P + 3, Q + 3, 16), w_code).float().permute(0, 3, 1, 2).contiguous()
        wt = _tt16(w, (K, 4, 4, 16), w_code).float().permute(0, 3, 1, 2).contiguous()
        _epilogue(Fn.conv2d(ci, wt), scale, shift, 0, 0, relu, y, y_dtype, None)

    def conv2d_stem_rgb(self, x, x_code, w, w_code, scale, shift, y, y_dtype, N, H, W, K, relu, code, st):
        import torch.nn.functional as Fn

        xi = _tt16(x, (N, H, W, 3), x_code).float().permute(0, 3, 1, 2).contiguous()
        wt = _tt16(w, (K, 192), w_code).float()[:, :147].reshape(K, 7, 7, 3).permute(0, 3, 1, 2).contiguous()
        _epilogue(Fn.conv2d(xi, wt, stride=2, padding=3), scale, shift, 0, 0, relu, y, y_dtype, None)

    def rgb_center(self, x, x_code, y, m0, m1, m2, pixels, st):
        import torch

        _tt(y, (pixels, 3)).copy_(_tt16(x, (pixels, 3), x_code).float() - torch.tensor([m0, m1, m2]))

    def im2col_depth_stem(self, x, a, a_code, N, H, W, st):
        import torch.nn.functional as Fn

        d = Fn.avg_pool2d(_tt(x, (N, H, W, 1)).permute(0, 3, 1, 2), 2)                # (N,1,H/2,W/2)
        cols = Fn.unfold(d, kernel_size=7, padding=3, stride=2)                         # (N, 49, P*Q), K index r*7+s
        out = _tt16(a, (N * cols.shape[2], 64), a_code)
        out.zero_()
        out[:, :49] = cols.permute(0, 2, 1).reshape(-1, 49).to(out.dtype)

    def im2col_rgb_stem(self, x, x_code, a, a_code, N, H, W, st):
        import torch.nn.functional as Fn

        img = _tt16(x, (N, H, W, 3), x_code).float().permute(0, 3, 1, 2)
        cols = Fn.unfold(img, kernel_size=7, padding=3, stride=2)                       # (N, 3*49, PQ): index c*49+r*7+s
        PQ = cols.shape[2]
        cols = cols.view(N, 3, 49, PQ).permute(0, 3, 2, 1).reshape(N * PQ, 147)        # -> (r*7+s)*3+c
        out = _tt16(a, (N * PQ, 192), a_code)
        out.zero_()
        out[:, :147] = cols.to(out.dtype)


def _wrap16(name, idx_codes):
    """re-dispatch an fp32-only model above when any dtype argument is 16-bit"""
    return None


# 16-bit aware versions of the normalisation / pooling models (the fp32 ones above stay as the strict checks)
def _groupnorm_fused(self, x, gamma, beta, residual, y, y_dtype, ws, wsb, sc, sh, N, HW, C, groups, eps, relu, st):
    import torch.nn.functional as Fn

    xi = _tt(x, (N, HW, C)).double().permute(0, 2, 1)
    out = Fn.group_norm(xi, groups, _tt(gamma, (C,)).double(), _tt(beta, (C,)).double(), eps)
    if residual:
        out = out + _tt16(residual, (N, HW, C), y_dtype).double().permute(0, 2, 1)
    if relu & 1:
        out = out.clamp_min(0)
    yt = _tt16(y, (N, HW, C), y_dtype)
    yt.copy_(out.permute(0, 2, 1).to(yt.dtype))


def _groupnorm_fused16(self, x, gamma, beta, residual, y, ws, wsb, N, HW, C, groups, eps, relu, st):
    import torch.nn.functional as Fn

    xi = _tt16(x, (N, HW, C), 4).double().permute(0, 2, 1)
    out = Fn.group_norm(xi, groups, _tt(gamma, (C,)).double(), _tt(beta, (C,)).double(), eps)
    if residual:
        out = out + _tt16(residual, (N, HW, C), 4).double().permute(0, 2, 1)
    if relu:
        out = out.clamp_min(0)
    yt = _tt16(y, (N, HW, C), 4)
    yt.copy_(out.permute(0, 2, 1).to(yt.dtype))


def _maxpool3x3s2(self, x, y, dtype, N, H, W, C, st):
    import torch.nn.functional as Fn

    P, Q = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = Fn.max_pool2d(_tt16(x, (N, H, W, C), dtype).float().permute(0, 3, 1, 2), 3, 2, 1)
    yt = _tt16(y, (N, P, Q, C), dtype)
    yt.copy_(out.permute(0, 2, 3, 1).to(yt.dtype))


def _adaptive_avgpool(self, x, dtype, y, ldy, N, H, W, C, O, st):
    import torch.nn.functional as Fn

    out = Fn.adaptive_avg_pool2d(_tt16(x, (N, H, W, C), dtype).double().permute(0, 3, 1, 2), (O, O))
    _ten3(y, N, O * O, C, O * O * ldy, ldy)[...] = out.flatten(2).permute(0, 2, 1).float().numpy()


def _affine_apply(self, x, x_dtype, scale, shift, scale_stride_n, residual, y, y_dtype, N, HW, C, relu, st):
    xi = _tt16(x, (N, HW, C), x_dtype).double()
    shp = (N, 1, C) if scale_stride_n else (1, 1, C)
    out = xi * _tt(scale, shp).double() + _tt(shift, shp).double()
    if residual:
        out = out + _tt16(residual, (N, HW, C), y_dtype).double()
    if relu:
        out = out.clamp_min(0)
    yt = _tt16(y, (N, HW, C), y_dtype)
    yt.copy_(out.to(yt.dtype))


def _batchnorm_train_stats(self, x, dtype, gamma, beta, rmean, rvar, momentum, scale, shift, ws, rows, C, eps, st):
    import torch

    xi = _tt16(x, (rows, C), dtype).double()
    mean, var = xi.mean(0), xi.var(0, unbiased=False)
    s = _tt(gamma, (C,)).double() / torch.sqrt(var + eps)
    _tt(scale, (C,)).copy_(s.float())
    _tt(shift, (C,)).copy_((_tt(beta, (C,)).double() - mean * s).float())
    rm, rv = _tt(rmean, (C,)), _tt(rvar, (C,))
    rm.copy_(((1 - momentum) * rm.double() + momentum * mean).float())
    rv.copy_(((1 - momentum) * rv.double() + momentum * xi.var(0, unbiased=True)).float())


for _name in [n for n in vars(TensorCoreOps) if not n.startswith("_")]:
    setattr(Emulator, _name, getattr(TensorCoreOps, _name))
Emulator.groupnorm_fused = _groupnorm_fused
Emulator.groupnorm_fused16 = _groupnorm_fused16
Emulator.maxpool3x3s2 = _maxpool3x3s2
Emulator.adaptive_avgpool = _adaptive_avgpool
Emulator.affine_apply = _affine_apply
Emulator.batchnorm_train_stats = _batchnorm_train_stats


# ======================================================================================
# device-resident trajectory cache (vlnce_b200/feature_cache.py, csrc/cache_ops.cu)
# ======================================================================================
_CACHE_NP = {0: np.float32, 2: np.uint8, 3: np.int64, 4: np.float16}


def _collate_rows(self, arena, dtype, offsets, lengths, out, T, N, D, pad, st):
    offs, lens = _flat(offsets, N, I64), _flat(lengths, N, I32)
    o = _flat(out, T * N * D).reshape(T, N, D)
    o[...] = pad
    for n in range(N):
        ln = min(int(lens[n]), T)
        if dtype == 1:  # bf16
            import torch

            src = _tt16(int(arena) + 2 * int(offs[n]), (ln, D), 1).float().numpy()
        else:
            nt = _CACHE_NP[dtype]
            ct = np.ctypeslib.as_ctypes_type(np.uint16 if nt is np.float16 else nt)
            raw = np.ctypeslib.as_array(ctypes.cast(int(arena) + int(offs[n]) * np.dtype(nt).itemsize, ctypes.POINTER(ct)),
                                        shape=(ln * D,))
            src = raw.view(nt).reshape(ln, D).astype(np.float32)
        o[:ln, n] = src


def _collate_actions(self, prev_arena, act_arena, offsets, lengths, prev_out, act_out, weights, masks, T, N, coef, st):
    offs, lens = _flat(offsets, N, I64), _flat(lengths, N, I32)
    po, ao = _flat(prev_out, T * N, I64).reshape(T, N), _flat(act_out, T * N, I64).reshape(T, N)
    w = _flat(weights, T * N).reshape(T, N)
    m = np.ctypeslib.as_array(ctypes.cast(int(masks), ctypes.POINTER(ctypes.c_uint8)), shape=(T * N,)).reshape(T, N)
    po[...] = 0
    ao[...] = 0
    w[...] = 0
    m[...] = 1
    m[0] = 0
    for n in range(N):
        ln = int(lens[n])
        a = _flat(int(act_arena) + 8 * int(offs[n]), ln, I64)
        po[:ln, n] = _flat(int(prev_arena) + 8 * int(offs[n]), ln, I64)
        ao[:ln, n] = a
        infl = np.concatenate([[True], a[1:] != a[:-1]])
        w[:ln, n] = np.where(infl, np.float32(coef), np.float32(1.0))


Emulator.collate_rows = _collate_rows
Emulator.collate_actions = _collate_actions


# ======================================================================================
# backward of the conv trunks (csrc/trunk_bwd.cu): torch-CPU functional models
# ======================================================================================
def _with_grad(fn):
    def wrapped(*a, **k):
        import torch

        with torch.enable_grad():
            return fn(*a, **k)

    wrapped.__name__ = fn.__name__
    return wrapped


class TrunkBwdOps:
    def conv2d_ref_dgrad(self, dy, w, dx, N, H, W, C, K, R, S, stride, pad, accumulate, st):
        import torch

        P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        g = _tt(dy, (N, P, Q, K)).permute(0, 3, 1, 2).double()
        wt = _tt(w, (K, R, S, C)).permute(0, 3, 1, 2).double()
        out = torch.nn.grad.conv2d_input((N, C, H, W), wt, g, stride=stride, padding=pad).permute(0, 2, 3, 1).float()
        t = _tt(dx, (N, H, W, C))
        t.copy_(t + out if accumulate else out)

    def conv2d_ref_wgrad(self, x, x_code, dy, dw, N, H, W, C, K, R, S, stride, pad, alpha, accumulate, st):
        import torch

        P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
        xi = _tt16(x, (N, H, W, C), x_code).double().permute(0, 3, 1, 2)
        g = _tt(dy, (N, P, Q, K)).permute(0, 3, 1, 2).double()
        out = torch.nn.grad.conv2d_weight(xi, (K, C, R, S), g, stride=stride, padding=pad).permute(0, 2, 3, 1)
        out = (alpha * out).float()
        t = _tt(dw, (K, R, S, C))
        t.copy_(t + out if accumulate else out)

    @_with_grad
    def groupnorm_bwd(self, x, dy, gamma, dx, partial, N, HW, C, groups, eps, st):
        import torch

        xi = _tt(x, (N, HW, C)).double().permute(0, 2, 1).contiguous().requires_grad_()
        gm = _tt(gamma, (C,)).double()
        g = _tt(dy, (N, HW, C)).double().permute(0, 2, 1)
        xr = xi.reshape(N, groups, -1)
        xhat = ((xr - xr.mean(2, keepdim=True)) / torch.sqrt(xr.var(2, unbiased=False, keepdim=True) + eps)).view(N, C, HW)
        (xhat * gm.view(1, C, 1) * g).sum().backward()
        _tt(dx, (N, HW, C)).copy_(xi.grad.permute(0, 2, 1).float())
        part = _tt(partial, (N, 2 * C))
        part[:, :C] = (g * xhat.detach()).sum(2).float()
        part[:, C:] = g.sum(2).float()

    @_with_grad
    def batchnorm_bwd(self, x, dy, gamma, rmean, rvar, dx, dgamma, dbeta, rows, C, eps, training, st):
        import torch

        xi = _tt(x, (rows, C)).double().clone().requires_grad_()
        g = _tt(dy, (rows, C)).double()
        if training:
            mean, var = xi.mean(0), xi.var(0, unbiased=False)
        else:
            mean, var = _tt(rmean, (C,)).double(), _tt(rvar, (C,)).double()
        xhat = (xi - mean) / torch.sqrt(var + eps)
        (xhat * _tt(gamma, (C,)).double() * g).sum().backward()
        _tt(dx, (rows, C)).copy_(xi.grad.float())
        _tt(dgamma, (C,)).copy_((g * xhat.detach()).sum(0).float())
        _tt(dbeta, (C,)).copy_(g.sum(0).float())

    @_with_grad
    def maxpool3x3s2_bwd(self, x, dy, dx, N, H, W, C, st):
        import torch.nn.functional as Fn

        P, Q = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        xi = _tt(x, (N, H, W, C)).permute(0, 3, 1, 2).clone().requires_grad_()
        Fn.max_pool2d(xi, 3, 2, 1).backward(_tt(dy, (N, P, Q, C)).permute(0, 3, 1, 2))
        _tt(dx, (N, H, W, C)).copy_(xi.grad.permute(0, 2, 3, 1))

    @_with_grad
    def adaptive_avgpool_bwd(self, dout, ld, dx, N, H, W, C, O, st):
        import torch
        import torch.nn.functional as Fn

        xi = torch.zeros(N, C, H, W, dtype=torch.float64, requires_grad=True)
        g = torch.from_numpy(_ten3(dout, N, O * O, C, O * O * ld, ld).copy()).double()     # (N, OO, C)
        Fn.adaptive_avg_pool2d(xi, (O, O)).backward(g.permute(0, 2, 1).reshape(N, C, O, O))
        _tt(dx, (N, H, W, C)).copy_(xi.grad.permute(0, 2, 3, 1).float())


for _name in [n for n in vars(TrunkBwdOps) if not n.startswith("_")]:
    setattr(Emulator, _name, getattr(TrunkBwdOps, _name))
