"""Hand-written backward of the CMA head (launch sequencing only; arithmetic in the library)
and the ``torch.autograd.Function`` that exposes it, so that the reference's
``loss.backward()`` (``common/base_il_trainer.py:172``) differentiates through
``policy.build_distribution(...).logits`` unchanged.  Only the trainable head has a
backward: both conv trunks are frozen in every shipped recipe (``config/default.py:242,253``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import lib
from .engine import linear_tc, F4, Branch, CMAEngine, gemm_group, linear, linear_dw, linear_dx, off
from .lib import C, ptr


def trainable_named(net, head_linear=None) -> Dict[str, torch.nn.Parameter]:
    out = {"net." + k: p for k, p in net.named_parameters() if p.requires_grad}
    if head_linear is not None:
        out["action_distribution.linear.weight"] = head_linear.weight
        out["action_distribution.linear.bias"] = head_linear.bias
    return out


def _colsum(a, lda, out, m, n, acc=0):
    from .engine import colsum

    colsum(a, lda, out, m, n, acc)


class _Leaves:
    """Weight/bias gradients are leaves of the backward graph.  ``defer=False``: each ``with leaves():`` block runs
    on a side stream while the dX chain continues on the caller's stream; ``join()`` once at the end.
    ``defer=True``: the blocks only QUEUE their GEMMs / column sums and ``join()`` issues all of them as one grouped
    launch (``vlnce_sgemm_grouped``) — their operands (saved activations and finished dY buffers) are not written
    again before the end of the backward, which the caller guarantees by opting in."""

    def __init__(self, defer: bool = False):
        from . import engine

        self.last = None
        self.defer = defer and engine.GROUPING and engine.DEFER_LEAVES
        self.pending = []

    def __call__(self):
        if self.defer:
            return _Queue(self.pending)
        self.last = Branch("dw")
        return self.last

    def join(self):
        if self.defer:
            from . import engine

            problems, self.pending = self.pending, []
            engine._flush(problems)
        elif self.last is not None:
            self.last.join()


class _Queue:
    """``with`` block whose GEMMs are appended to ``sink`` instead of being launched."""

    def __init__(self, sink):
        self.sink = sink

    def __enter__(self):
        from . import engine

        self.outer = engine._PENDING  # (an open gemm_group keeps collecting after this block)
        engine._PENDING = self.sink
        return self

    def __exit__(self, *exc):
        from . import engine

        engine._PENDING = self.outer
        return False


def _gru_backward(eng: CMAEngine, tag, rnn, pfx, g, dh, lddh, x, ldx, in_size, masks_f, T, N, dx, lddx, leaves):
    """BPTT of ``HeadEngine._gru_forward`` (GRU, or LSTM when the state encoder is one)."""
    if isinstance(rnn, torch.nn.LSTM):
        return _lstm_backward(eng, tag, rnn, pfx, g, dh, lddh, x, ldx, in_size, masks_f, T, N, dx, lddx, leaves)
    H = eng.H
    R = T * N
    st = lib.stream()
    d_gi = eng.buf(tag + "_dgi", (R, 3 * H))
    d_gh = eng.buf(tag + "_dgh", (R, 3 * H))
    dhm = eng.buf(tag + "_dhm", (N, H))
    save = eng.buf(tag + "_save", (R, 4 * H))
    hm_all = eng.buf(tag + "_hm", (R, H))
    for t in range(T - 1, -1, -1):
        C.vlnce_gru_pointwise_bwd(dh + F4 * t * N * lddh, lddh, off(save, t * N * 4 * H), off(hm_all, t * N * H),
                                  off(d_gi, t * N * 3 * H), off(d_gh, t * N * 3 * H), ptr(dhm), N, H, st)
        if t > 0:  # (the gradient w.r.t. the initial hidden state has no consumer: rnn_states are inputs)
            linear_dx(off(d_gh, t * N * 3 * H), 3 * H, ptr(rnn.weight_hh_l0), ptr(dhm), H, N, 3 * H, H, beta=1.0)
            C.vlnce_mask_rows_add(ptr(dhm), H, off(masks_f, t * N), dh + F4 * (t - 1) * N * lddh, lddh, N, H, st)
    with leaves():
        linear_dw(ptr(d_gh), 3 * H, ptr(hm_all), H, ptr(g[pfx + "weight_hh_l0"]), H, R, 3 * H, H)
        _colsum(ptr(d_gh), 3 * H, ptr(g[pfx + "bias_hh_l0"]), R, 3 * H)
        linear_dw(ptr(d_gi), 3 * H, x, ldx, ptr(g[pfx + "weight_ih_l0"]), in_size, R, 3 * H, in_size)
        _colsum(ptr(d_gi), 3 * H, ptr(g[pfx + "bias_ih_l0"]), R, 3 * H)
    linear_dx(ptr(d_gi), 3 * H, ptr(rnn.weight_ih_l0), dx, lddx, R, 3 * H, in_size)


def _lstm_backward(eng, tag, rnn, pfx, g, dh, lddh, x, ldx, in_size, masks_f, T, N, dx, lddx, leaves):
    """LSTM state encoder: gates = gi + gh, so one pre-activation gradient buffer feeds both weight matrices;
    the cell-state gradient is carried (and mask-reset) between steps inside the pointwise kernel."""
    H = eng.H
    R = T * N
    st = lib.stream()
    d_g = eng.buf(tag + "_dgi", (R, 4 * H))
    dhm = eng.buf(tag + "_dhm", (N, H))
    dc = eng.buf(tag + "_dc", (N, H))
    save = eng.buf(tag + "_save", (R, 5 * H))
    hm_all, cm_all = eng.buf(tag + "_hm", (R, H)), eng.buf(tag + "_cm", (R, H))
    for t in range(T - 1, -1, -1):
        C.vlnce_lstm_pointwise_bwd(dh + F4 * t * N * lddh, lddh, ptr(dc) if t < T - 1 else None,
                                   off(save, t * N * 5 * H), off(cm_all, t * N * H), off(masks_f, t * N),
                                   off(d_g, t * N * 4 * H), ptr(dc), N, H, st)
        if t > 0:
            linear_dx(off(d_g, t * N * 4 * H), 4 * H, ptr(rnn.weight_hh_l0), ptr(dhm), H, N, 4 * H, H)
            C.vlnce_mask_rows_add(ptr(dhm), H, off(masks_f, t * N), dh + F4 * (t - 1) * N * lddh, lddh, N, H, st)
    with leaves():
        linear_dw(ptr(d_g), 4 * H, ptr(hm_all), H, ptr(g[pfx + "weight_hh_l0"]), H, R, 4 * H, H)
        _colsum(ptr(d_g), 4 * H, ptr(g[pfx + "bias_hh_l0"]), R, 4 * H)
        linear_dw(ptr(d_g), 4 * H, x, ldx, ptr(g[pfx + "weight_ih_l0"]), in_size, R, 4 * H, in_size)
        _colsum(ptr(d_g), 4 * H, ptr(g[pfx + "bias_ih_l0"]), R, 4 * H)
    linear_dx(ptr(d_g), 4 * H, ptr(rnn.weight_ih_l0), dx, lddx, R, 4 * H, in_size)


def instruction_backward(eng, g: Dict[str, torch.Tensor], d_instr: torch.Tensor):
    """BPTT of the packed (bi)LSTM / (bi)GRU instruction encoder (instruction_encoder.py:63-94) from the
    gradient w.r.t. its per-token outputs ``d_instr`` (R, L, dirs*Hs): pre-activation gate gradients
    from the persistent kernel, then the weight / embedding gradients as GEMMs + a scatter."""
    net, cx, b = eng.net, eng.ctx, eng.buf
    R, L = cx["R"], cx["L"]
    ie = net.instruction_encoder
    rnn = ie.encoder_rnn
    Hs, dirs, E = eng.Hs, eng.dirs, eng.E
    instr = b("instr_out", (R, L, dirs * Hs))
    gates = 3 if isinstance(rnn, torch.nn.GRU) else 4
    G = gates * Hs
    dgx = b("lstm_dgx", (dirs, R, L, G))
    hprev = b("lstm_hprev", (dirs, R, L, Hs))
    whh_b = ptr(rnn.weight_hh_l0_reverse) if dirs == 2 else None
    if gates == 3:  # GRU: the hidden-side pre-activation gradient differs from the input-side one (candidate rows * r)
        dgh = b("gru_seq_dgh", (dirs, R, L, G))
        C.vlnce_gru_seq_bwd(ptr(d_instr), ptr(b("gru_seq_save", (dirs, R, L, 4 * Hs))), ptr(instr),
                            ptr(rnn.weight_hh_l0), whh_b, ptr(cx["lengths"]), ptr(dgx), ptr(dgh), ptr(hprev), R, L,
                            Hs, dirs, lib.stream())
    else:
        dgh = dgx
        C.vlnce_lstm_seq_bwd(ptr(d_instr), ptr(b("lstm_gates", (dirs, R, L, 4 * Hs))), ptr(b("lstm_c", (dirs, R, L, Hs))),
                             ptr(instr), ptr(rnn.weight_hh_l0), whh_b, ptr(cx["lengths"]), ptr(dgx), ptr(hprev), R, L,
                             Hs, dirs, lib.stream())
    emb_in = b("instr_emb", (R * L, E)) if ie.config.sensor_uuid == "instruction" else cx["rxr"]
    emb_trainable = ie.config.sensor_uuid == "instruction" and ie.embedding_layer.weight.requires_grad
    d_emb = b("d_instr_emb", (R * L, E)) if emb_trainable else None
    pfx = "net.instruction_encoder.encoder_rnn."
    M = R * L
    tc = bool(cx.get("instr_tc")) and M % 32 == 0 and E % 64 == 0
    if tc:
        # weight gradient of the input projection on the tensor cores: dW[d] = dgx[d]^T @ features is the TMA tile's GEMM
        # with the token axis as the reduction dimension — both operands transposed to token-minor (and rounded to TF32)
        # first; the directions are stacked into ONE (dirs*G, M) x (E, M)^T launch
        emb_t = b("instr_emb_t_tf32", (E, M))
        C.vlnce_transpose_f32(ptr(b("instr_emb_tf32", (M, E))), E, ptr(emb_t), M, M, E, 0, lib.stream())
        dgx_t = b("lstm_dgx_t_tf32", (dirs * G, M))
        for d in range(dirs):
            C.vlnce_transpose_f32(off(dgx, d * M * G), G, off(dgx_t, d * G * M), M, M, G, 1, lib.stream())
        dw_all = b("instr_dwih", (dirs * G, E))
        linear_tc(ptr(dgx_t), ptr(emb_t), ptr(dw_all), None, dirs * G, E, M, dgx.device)
        for d, sfx in enumerate(["", "_reverse"][:dirs]):
            g[pfx + "weight_ih_l0" + sfx].copy_(dw_all[d * G:(d + 1) * G])
    with gemm_group():  # every problem of this block is independent of the others: one launch
        for d, sfx in enumerate(["", "_reverse"][:dirs]):
            dg, dh_ = off(dgx, d * R * L * G), off(dgh, d * R * L * G)
            if not tc:
                linear_dw(dg, G, ptr(emb_in), E, ptr(g[pfx + "weight_ih_l0" + sfx]), E, R * L, G, E)
            linear_dw(dh_, G, off(hprev, d * R * L * Hs), Hs, ptr(g[pfx + "weight_hh_l0" + sfx]), Hs, R * L, G, Hs)
            _colsum(dg, G, ptr(g[pfx + "bias_ih_l0" + sfx]), R * L, G)
            _colsum(dh_, G, ptr(g[pfx + "bias_hh_l0" + sfx]), R * L, G)
        if emb_trainable:
            linear_dx(ptr(dgx), G, ptr(rnn.weight_ih_l0), ptr(d_emb), E, R * L, G, E)
    if emb_trainable and dirs == 2:  # accumulates into the forward direction's d_emb: after the group
        linear_dx(off(dgx, R * L * G), G, ptr(rnn.weight_ih_l0_reverse), ptr(d_emb), E, R * L, G, E, beta=1.0)
    if emb_trainable:
        el = ie.embedding_layer
        pad = -1 if el.padding_idx is None else el.padding_idx
        C.vlnce_embedding_bwd(ptr(cx["ids"]), ptr(d_emb), E, ptr(g["net.instruction_encoder.embedding_layer.weight"]),
                              R * L, E, el.num_embeddings, pad, lib.stream())


def head_and_progress_backward(eng, head_linear, g, d_raw, d_x, d_hat, leaves):
    """Steps shared by both nets: action head (or a feature gradient) and the progress monitor
    ``hat = tanh(x w + b)``; returns the gradient buffer w.r.t. the net's output features."""
    net, cx, b = eng.net, eng.ctx, eng.buf
    R, H = cx["R"], eng.H
    st = lib.stream()
    x_out = b("x_out", (R, H))
    dx_out = b("d_x_out", (R, H))
    if head_linear is not None:
        A = head_linear.weight.shape[0]
        with leaves():
            linear_dw(ptr(d_raw), A, ptr(x_out), H, ptr(g["action_distribution.linear.weight"]), H, R, A, H)
            _colsum(ptr(d_raw), A, ptr(g["action_distribution.linear.bias"]), R, A)
        linear_dx(ptr(d_raw), A, ptr(head_linear.weight), ptr(dx_out), H, R, A, H)
    else:
        dx_out.copy_(d_x)
    if d_hat is not None and "net.progress_monitor.weight" in g:
        hat = b("progress_hat", (R, 1))
        dpre = b("d_pm", (R, 1))
        dpre.copy_(d_hat.reshape(R, 1))
        C.vlnce_tanh_bwd(ptr(dpre), ptr(hat), R, st)
        pm = net.progress_monitor
        with leaves():
            linear_dw(ptr(dpre), 1, ptr(x_out), H, ptr(g["net.progress_monitor.weight"]), H, R, 1, H)
            _colsum(ptr(dpre), 1, ptr(g["net.progress_monitor.bias"]), R, 1)
        linear_dx(ptr(dpre), 1, ptr(pm.weight), ptr(dx_out), H, R, 1, H, beta=1.0)
    return dx_out


def seq2seq_backward(eng, head_linear, g: Dict[str, torch.Tensor], d_raw: Optional[torch.Tensor],
                     d_x: Optional[torch.Tensor], d_hat: Optional[torch.Tensor]):
    """Backward of ``Seq2SeqEngine.forward(..., save=True)`` (seq2seq_policy.py:142-179): action head /
    progress monitor -> GRU (BPTT over the T rows blocks) -> the four concat slices: instruction final
    state (scattered to its token, then the LSTM's BPTT), depth fc, rgb fc (ReLU(Linear); the conv
    trunks are frozen, ``config/default.py:242,253``) and the optional prev-action embedding."""
    net, cx, b = eng.net, eng.ctx, eng.buf
    assert cx.get("saved"), "backward needs the forward to have run with save=True"
    R, N, T, L = cx["R"], cx["N"], cx["T"], cx["L"]
    H, Hs = eng.H, eng.Hs
    d_out, r_out, width = cx["d_out"], cx["r_out"], cx["width"]
    st = lib.stream()
    masks_f = b("masks_f", (R,))
    xin = b("s2s_in", (R, width))
    leaves = _Leaves(defer=True)  # operands of every leaf stay untouched until the end of this backward
    dx_out = head_and_progress_backward(eng, head_linear, g, d_raw, d_x, d_hat, leaves)
    d_xin = b("d_s2s_in", (R, width))
    _gru_backward(eng, "gru", net.state_encoder.rnn, "net.state_encoder.rnn.", g, ptr(dx_out), H, ptr(xin), width,
                  width, masks_f, T, N, ptr(d_xin), width, leaves)
    # instruction final state -> its token's output gradient -> LSTM BPTT (own stream)
    with Branch("lstm_bwd") as br_lstm:
        if not net.model_config.get("ablate_instruction", False):  # ``instruction_embedding * 0`` stops the gradient
            d_instr = b("d_instr", (R, L, Hs))
            C.vlnce_scatter_last_token(ptr(d_xin), width, ptr(cx["lengths"]), ptr(d_instr), R, L, Hs, lib.stream())
            instruction_backward(eng, g, d_instr)
    # depth fc | rgb fc: ReLU'(y) then dW, db (inputs saved by the forward).  An ablated slice was cleared in the
    # forward, so ReLU' = 0 there and its fc gradients come out exactly 0, as in the reference.
    C.vlnce_relu_bwd(off(d_xin, Hs), width, off(xin, Hs), width, R, d_out + r_out, st)
    c, h, w = net.depth_encoder.visual_encoder.output_shape
    dep_flat = b("s2s_depth_flat", (R, c * h * w))
    if "net.depth_encoder.visual_fc.1.weight" in g:
        with leaves():
            linear_dw(off(d_xin, Hs), width, ptr(dep_flat), c * h * w, ptr(g["net.depth_encoder.visual_fc.1.weight"]),
                      c * h * w, R, d_out, c * h * w)
            _colsum(off(d_xin, Hs), width, ptr(g["net.depth_encoder.visual_fc.1.bias"]), R, d_out)
    cr = net.rgb_encoder.resnet_layer_size
    rgb = b("s2s_rgb", (R, 1, cr))
    if "net.rgb_encoder.fc.1.weight" in g:
        with leaves():
            linear_dw(off(d_xin, Hs + d_out), width, ptr(rgb), cr, ptr(g["net.rgb_encoder.fc.1.weight"]), cr, R,
                      r_out, cr)
            _colsum(off(d_xin, Hs + d_out), width, ptr(g["net.rgb_encoder.fc.1.bias"]), R, r_out)
    if net.model_config.SEQ2SEQ.use_prev_action:  # seq2seq_policy.py:158-162
        pa_w = net.prev_action_embedding.weight
        C.vlnce_prev_action_embed_bwd(ptr(cx["prev"]), ptr(masks_f), off(d_xin, width - 32), width, None, 0,
                                      ptr(g["net.prev_action_embedding.weight"]), R, 32, pa_w.shape[0], st)
    # trainable conv trunks (seq2seq_policy.py:78,92): fc input gradients -> the trunks' taped backward.  An ablated
    # modality's slice has zero gradient after the ReLU mask above, as in the reference (``embedding * 0``).
    taped = [e for e in (net.rgb_encoder, net.depth_encoder) if getattr(e, "_taped", None) is not None]
    pname = {id(p): k for k, p in trainable_named(net).items()} if taped else {}
    if getattr(net.depth_encoder, "_taped", None) is not None:
        P = h * w
        d_flat = b("d_s2s_depth_flat", (R, c * P))
        linear_dx(off(d_xin, Hs), width, ptr(net.depth_encoder.visual_fc[1].weight), ptr(d_flat), c * P, R, d_out, c * P)
        d_tr = b("d_s2s_depth", (R, P, c))
        C.vlnce_nchw_to_positions(ptr(d_flat), ptr(d_tr), c, R, c, P, st)
        net.depth_encoder.backward_trunk(ptr(d_tr), c, g, pname)
    if getattr(net.rgb_encoder, "_taped", None) is not None:
        d_rgb = b("d_s2s_rgb", (R, 1, cr))
        linear_dx(off(d_xin, Hs + d_out), width, ptr(net.rgb_encoder.fc[1].weight), ptr(d_rgb), cr, R, r_out, cr)
        net.rgb_encoder.backward_trunk(ptr(d_rgb), cr, g, pname)
    br_lstm.join()
    leaves.join()


def cma_backward(eng: CMAEngine, head_linear, g: Dict[str, torch.Tensor], d_raw: Optional[torch.Tensor],
                 d_x: Optional[torch.Tensor], d_hat: Optional[torch.Tensor]):
    """Backward of the last ``eng.forward(..., save=True)``.  ``g`` maps parameter names to
    ZEROED gradient tensors that are written here.  ``d_raw`` (R,A): gradient w.r.t. the raw
    action logits (with ``head_linear``) or ``d_x`` (R,H) w.r.t. the features; ``d_hat`` (R,1)
    w.r.t. ``progress_hat``.  All buffers are the engine's persistent workspace."""
    net = eng.net
    cx = eng.ctx
    assert cx.get("saved"), "backward needs the forward to have run with save=True"
    R, N, T, L = cx["R"], cx["N"], cx["T"], cx["L"]
    H, P, hk = eng.H, eng.P, eng.H // 2
    st = lib.stream()
    b = eng.buf
    x_out, x2, xcat, state_in = b("x_out", (R, H)), b("x2", (R, H)), b("xcat", (R, eng.cat)), b("state_in", (R, eng.state_in))
    masks_f = b("masks_f", (R,))
    dx_out = b("d_x_out", (R, H))
    # the weight / bias gradients are leaves whose operands stay untouched until the end of this backward (saved
    # forward buffers and finished dY buffers, each with its own name): all of them go out as ONE grouped launch
    leaves = _Leaves(defer=True)

    # 1. action head / feature gradient
    if head_linear is not None:
        A = head_linear.weight.shape[0]
        with leaves():
            linear_dw(ptr(d_raw), A, ptr(x_out), H, ptr(g["action_distribution.linear.weight"]), H, R, A, H)
            _colsum(ptr(d_raw), A, ptr(g["action_distribution.linear.bias"]), R, A)
        linear_dx(ptr(d_raw), A, ptr(head_linear.weight), ptr(dx_out), H, R, A, H)
    else:
        dx_out.copy_(d_x)
    # 2. progress monitor: hat = tanh(x w + b)
    if d_hat is not None and "net.progress_monitor.weight" in g:
        hat = b("progress_hat", (R, 1))
        dpre = b("d_pm", (R, 1))
        dpre.copy_(d_hat.reshape(R, 1))
        C.vlnce_tanh_bwd(ptr(dpre), ptr(hat), R, st)
        pm = net.progress_monitor
        with leaves():
            linear_dw(ptr(dpre), 1, ptr(x_out), H, ptr(g["net.progress_monitor.weight"]), H, R, 1, H)
            _colsum(ptr(dpre), 1, ptr(g["net.progress_monitor.bias"]), R, 1)
        linear_dx(ptr(dpre), 1, ptr(pm.weight), ptr(dx_out), H, R, 1, H, beta=1.0)
    # 3. second GRU
    d_x2 = b("d_x2", (R, H))
    _gru_backward(eng, "gru2", net.second_state_encoder.rnn, "net.second_state_encoder.rnn.", g, ptr(dx_out), H,
                  ptr(x2), H, H, masks_f, T, N, ptr(d_x2), H, leaves)
    # 4. second_state_compress = ReLU(Linear(xcat))
    C.vlnce_relu_bwd(ptr(d_x2), H, ptr(x2), H, R, H, st)
    lin = net.second_state_compress[0]
    with leaves():
        linear_dw(ptr(d_x2), H, ptr(xcat), eng.cat, ptr(g["net.second_state_compress.0.weight"]), eng.cat, R, H, eng.cat)
        _colsum(ptr(d_x2), H, ptr(g["net.second_state_compress.0.bias"]), R, H)
    dxcat = b("d_xcat", (R, eng.cat))
    linear_dx(ptr(d_x2), H, ptr(lin.weight), ptr(dxcat), eng.cat, R, H, eng.cat)
    # 5. visual attentions
    text_q = b("text_q", (R, hk))
    d_text_q = b("d_text_q", (R, hk))
    dE = {}
    vis = (("rgb", eng.Crgb, net.rgb_kv, eng.rgb_out, H + eng.Dt),
           ("depth", eng.Cdep, net.depth_kv, eng.depth_out, H + eng.Dt + eng.rgb_out))
    for name, Cc, kv, v_out, col in vis:
        gw, gb = g[f"net.{name}_kv.weight"], g[f"net.{name}_kv.bias"]
        with leaves():
            linear_dw(off(dxcat, col), eng.cat, ptr(b("pooled_" + name, (R, Cc))), Cc, off(gw, hk * Cc), Cc, R, v_out,
                      Cc)  # dW_v
            _colsum(off(dxcat, col), eng.cat, off(gb, hk), R, v_out)
    with gemm_group():  # gradients w.r.t. the two pooled embeddings: independent, one launch
        for name, Cc, kv, v_out, col in vis:
            linear_dx(off(dxcat, col), eng.cat, off(kv.weight, hk * Cc), ptr(b("d_pooled_" + name, (R, Cc))), Cc, R,
                      v_out, Cc)
    for name, Cc, kv, v_out, col in vis:
        E_ = cx["rgb_e"] if name == "rgb" else cx["dep_e"]
        qv, probs = b("q_" + name, (R, Cc)), b("p_" + name, (R, P))
        dq = b("d_q_" + name, (R, Cc))
        dEn = b("d_emb_" + name, (R, P, Cc))
        C.vlnce_attn_pool_bwd(ptr(qv), Cc, ptr(E_), P * Cc, Cc, ptr(probs), eng.scale, ptr(b("d_pooled_" + name, (R, Cc))),
                              Cc, ptr(dq), Cc, ptr(dEn), P * Cc, Cc, 0, R, P, Cc, st)
        dE[name] = dEn
        with leaves():
            linear_dw(ptr(text_q), hk, ptr(dq), Cc, ptr(g[f"net.{name}_kv.weight"]), Cc, R, hk, Cc)  # dW_k = text_q^T dq'
    for i, (name, Cc, kv, v_out, col) in enumerate(vis):  # both accumulate into d_text_q: in order
        linear(ptr(b("d_q_" + name, (R, Cc))), Cc, ptr(kv.weight), ptr(d_text_q), hk, None, R, hk, Cc,
               beta=0.0 if i == 0 else 1.0)
    # 6. text_q = Linear(text_embedding)
    with leaves():
        linear_dw(ptr(d_text_q), hk, off(xcat, H), eng.cat, ptr(g["net.text_q.weight"]), eng.Dt, R, hk, eng.Dt)
        _colsum(ptr(d_text_q), hk, ptr(g["net.text_q.bias"]), R, hk)
    linear_dx(ptr(d_text_q), hk, ptr(net.text_q.weight), off(dxcat, H), eng.cat, R, hk, eng.Dt, beta=1.0)
    # 7. text attention
    instr = b("instr_out", (R, L, eng.Dt))
    q_t, q_s, p_text = b("q_text", (R, eng.Dt)), b("q_state", (R, hk)), b("p_text", (R, L))
    dq_t = b("d_q_text", (R, eng.Dt))
    d_instr = b("d_instr", (R, L, eng.Dt))
    C.vlnce_attn_pool_bwd(ptr(q_t), eng.Dt, ptr(instr), L * eng.Dt, eng.Dt, ptr(p_text), eng.scale, off(dxcat, H),
                          eng.cat, ptr(dq_t), eng.Dt, ptr(d_instr), L * eng.Dt, eng.Dt, 0, R, L, eng.Dt, st)
    with leaves():
        linear_dw(ptr(q_s), hk, ptr(dq_t), eng.Dt, ptr(g["net.text_k.weight"]), eng.Dt, R, hk, eng.Dt)
    d_q_s = b("d_q_state", (R, hk))
    linear(ptr(dq_t), eng.Dt, ptr(net.text_k.weight), ptr(d_q_s), hk, None, R, hk, eng.Dt)
    with leaves():
        linear_dw(ptr(d_q_s), hk, ptr(xcat), eng.cat, ptr(g["net.state_q.weight"]), H, R, hk, H)
        _colsum(ptr(d_q_s), hk, ptr(g["net.state_q.bias"]), R, hk)
    linear_dx(ptr(d_q_s), hk, ptr(net.state_q.weight), ptr(dxcat), eng.cat, R, hk, H, beta=1.0)
    cfg = net.model_config
    with Branch("lstm_bwd") as br_lstm:
        # 12. instruction encoder (BPTT) — independent of steps 8-11: own stream
        if not cfg.get("ablate_instruction", False):  # ``instruction_embedding * 0``: nothing flows upstream
            instruction_backward(eng, g, d_instr)

    # 8. first GRU (dh = dxcat[:, :H])
    d_state_in = b("d_state_in", (R, eng.state_in))
    _gru_backward(eng, "gru1", net.state_encoder.rnn, "net.state_encoder.rnn.", g, ptr(dxcat), eng.cat,
                  ptr(state_in), eng.state_in, eng.state_in, masks_f, T, N, ptr(d_state_in), eng.state_in, leaves)
    # 9. rgb_linear / depth_linear (ReLU(Linear))
    c_rgb = net.rgb_encoder.resnet_layer_size
    c_dep = net.depth_encoder.visual_encoder.output_shape[0]
    C.vlnce_relu_bwd(ptr(d_state_in), eng.state_in, ptr(state_in), eng.state_in, R, eng.rgb_out + eng.depth_out, st)
    rgb_mean = b("rgb_mean", (R, eng.Crgb))
    lin = net.rgb_linear[2]
    with leaves():
        linear_dw(ptr(d_state_in), eng.state_in, ptr(rgb_mean), eng.Crgb, ptr(g["net.rgb_linear.2.weight"]), eng.Crgb,
                  R, eng.rgb_out, eng.Crgb)
        _colsum(ptr(d_state_in), eng.state_in, ptr(g["net.rgb_linear.2.bias"]), R, eng.rgb_out)
    d_mean_sp = b("d_mean_sp", (R, 64))
    d_flat_sp = b("d_flat_sp", (R, 64 * P))
    dep_flat = b("depth_flat", (R, eng.Cdep * P))
    lin_d = net.depth_linear[1]
    d_dep = off(d_state_in, eng.rgb_out)
    with leaves():
        linear_dw(d_dep, eng.state_in, ptr(dep_flat), eng.Cdep * P, ptr(g["net.depth_linear.1.weight"]), eng.Cdep * P,
                  R, eng.depth_out, eng.Cdep * P)
        _colsum(d_dep, eng.state_in, ptr(g["net.depth_linear.1.bias"]), R, eng.depth_out)
    with gemm_group():  # the two spatial-embedding input gradients are independent: one launch
        linear_dx(ptr(d_state_in), eng.state_in, off(lin.weight, c_rgb), ptr(d_mean_sp), 64, R, eng.rgb_out, 64,
                  ldw=eng.Crgb)
        linear_dx(d_dep, eng.state_in, off(lin_d.weight, c_dep * P), ptr(d_flat_sp), 64 * P, R, eng.depth_out, 64 * P,
                  ldw=eng.Cdep * P)
    # 10. prev-action embedding (two concat slices share the rows)
    pa_w = net.prev_action_embedding.weight
    C.vlnce_prev_action_embed_bwd(ptr(cx["prev"]), ptr(masks_f), off(dxcat, eng.cat - 32), eng.cat,
                                  off(d_state_in, eng.rgb_out + eng.depth_out), eng.state_in,
                                  ptr(g["net.prev_action_embedding.weight"]), R, 32, pa_w.shape[0], st)
    # 11. spatial embeddings
    if not cfg.get("ablate_rgb", False):
        C.vlnce_spatial_embed_grad(ptr(dE["rgb"]), P * eng.Crgb, eng.Crgb, c_rgb, None, 0, ptr(d_mean_sp), 64,
                                   ptr(g["net.rgb_encoder.spatial_embeddings.weight"]), 0, R, P, 64, st)
    if not cfg.get("ablate_depth", False):
        C.vlnce_spatial_embed_grad(ptr(dE["depth"]), P * eng.Cdep, eng.Cdep, c_dep, ptr(d_flat_sp), 64 * P, None, 0,
                                   ptr(g["net.depth_encoder.spatial_embeddings.weight"]), 0, R, P, 64, st)
    # 13. trainable conv trunks (RGB_ENCODER.trainable / DEPTH_ENCODER.trainable, cma_policy.py:80,94): the trunk features
    # feed the attention (dE) AND the rgb_linear / depth_linear inputs (steps 9), whose trunk-channel gradients were
    # skipped above; both are gathered into a contiguous (R, P, c) buffer the trunk's taped backward starts from
    taped = [e for e in (net.rgb_encoder, net.depth_encoder) if getattr(e, "_taped", None) is not None]
    pname = {id(p): k for k, p in trainable_named(net).items()} if taped else {}  # (walks the module tree: only when needed)
    if getattr(net.rgb_encoder, "_taped", None) is not None and not cfg.get("ablate_rgb", False):
        d_mean = b("d_rgb_mean_tr", (R, c_rgb))
        linear_dx(ptr(d_state_in), eng.state_in, ptr(net.rgb_linear[2].weight), ptr(d_mean), c_rgb, R, eng.rgb_out,
                  c_rgb, ldw=eng.Crgb)
        d_mean_p = b("d_rgb_mean_tr_p", (R, c_rgb))
        d_mean_p.zero_()
        C.vlnce_axpy(ptr(d_mean_p), ptr(d_mean), 1.0 / P, R * c_rgb, st)  # mean over the P positions
        d_tr = b("d_rgb_trunk", (R, P, c_rgb))
        C.vlnce_broadcast_rows(ptr(d_mean_p), c_rgb, ptr(d_tr), P * c_rgb, c_rgb, R, P, c_rgb, st)
        C.vlnce_sum_positions(ptr(dE["rgb"]), eng.Crgb, eng.Crgb, ptr(d_tr), c_rgb, 1, R * P, 1, c_rgb, st)
        net.rgb_encoder.backward_trunk(ptr(d_tr), c_rgb, g, pname)
    if getattr(net.depth_encoder, "_taped", None) is not None and not cfg.get("ablate_depth", False):
        d_flat = b("d_dep_flat_tr", (R, c_dep * P))  # NCHW-flat trunk channels of depth_linear's input
        linear_dx(d_dep, eng.state_in, ptr(net.depth_linear[1].weight), ptr(d_flat), c_dep * P, R, eng.depth_out,
                  c_dep * P, ldw=eng.Cdep * P)
        d_tr = b("d_dep_trunk", (R, P, c_dep))
        C.vlnce_nchw_to_positions(ptr(d_flat), ptr(d_tr), c_dep, R, c_dep, P, st)
        C.vlnce_sum_positions(ptr(dE["depth"]), eng.Cdep, eng.Cdep, ptr(d_tr), c_dep, 1, R * P, 1, c_dep, st)
        net.depth_encoder.backward_trunk(ptr(d_tr), c_dep, g, pname)
    br_lstm.join()
    leaves.join()


class CMAFunction(torch.autograd.Function):
    """(trainable parameters…) -> (features | log-prob logits, rnn_states_out, progress_hat)."""

    @staticmethod
    def run(net, head_linear, observations, rnn_states, prev_actions, masks):
        named = trainable_named(net, head_linear)
        out, h_out, hat = CMAFunction.apply(net, head_linear, observations, rnn_states, prev_actions, masks,
                                            tuple(named.keys()), *named.values())
        return out, h_out, (hat if hat.numel() else None)

    @staticmethod
    def forward(ctx, net, head_linear, observations, rnn_states, prev_actions, masks, names, *params):
        eng = net.engine()
        x, h_out, hat = eng.forward(observations, rnn_states, prev_actions, masks, save=True)
        ctx.eng, ctx.head, ctx.names = eng, head_linear, names
        ctx.shapes = [p.shape for p in params]
        ctx.token = object()
        eng.ctx["token"] = ctx.token
        out = x if head_linear is None else eng.action_logits(head_linear, x, x.shape[0])
        h_ret = h_out.clone()
        ctx.mark_non_differentiable(h_ret)
        hat_ret = hat.clone() if hat is not None else x.new_zeros(0)
        return out.clone(), h_ret, hat_ret

    @staticmethod
    def backward(ctx, d_out, _d_h, d_hat):
        eng = ctx.eng  # CMAEngine | Seq2SeqEngine: both expose run_backward
        if eng.ctx.get("token") is not ctx.token:
            raise lib.VlnceError("backward() after another forward on the same policy: the saved workspace was "
                                 "overwritten (the reference's trainers never interleave them)")
        sizes = [int(torch.Size(s).numel()) for s in ctx.shapes]
        flat = torch.zeros(sum(sizes), device=d_out.device, dtype=torch.float32)
        g, o = {}, 0
        for name, shp, n in zip(ctx.names, ctx.shapes, sizes):
            g[name] = flat[o:o + n].view(shp)
            o += n
        d_out = d_out.contiguous().float()
        d_hat = None if (d_hat is None or d_hat.numel() == 0) else d_hat.contiguous().float()
        if ctx.head is not None:
            R, A = d_out.shape
            lp = eng.buf("logits", (R, A))
            d_raw = eng.buf("d_raw", (R, A))
            C.vlnce_log_softmax_bwd(ptr(lp), ptr(d_out), ptr(d_raw), R, A, lib.stream())
            eng.run_backward(ctx.head, g, d_raw, None, d_hat)
        else:
            eng.run_backward(None, g, None, d_out, d_hat)
        return (None,) * 7 + tuple(g[n] for n in ctx.names)
