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
