"""Launch sequencing of the CMA head (forward and hand-written backward) over the C-ABI
kernels.  Follows ``CMANet.forward`` (``cma_policy.py:219-309``) step by step; the only
algebraic change is the re-association of the three single-query attentions
(``_attn``, ``cma_policy.py:207-217``): with k_p = W_k e_p + b_k,

    q.k_p = (W_k^T q).e_p + q.b_k        (the constant cancels in the softmax)
    sum_p a_p (W_v e_p + b_v) = W_v (sum_p a_p e_p) + b_v

so keys/values are never materialised — the raw embedding is attention-pooled once and the
value projection runs on the pooled vector (16x fewer MACs for the visual attentions, 80x for
the text keys).  Everything is fp32; rows are time-major (T*N, .) as ``collate_fn`` builds them.

Python here only orders kernel launches on the current CUDA stream and owns the persistent
workspace (static addresses, CUDA-graph capturable).  No tensor arithmetic happens in torch.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import lib
from .lib import C, ptr
from .trunk import BufferPool


_WS = {}


def gemm_workspace(nbytes: int = 16 << 20):
    """Caller-owned split-K scratch, one per (device, stream) — concurrent branches of the
    step run on their own streams and must not share scratch."""
    if torch.cuda.is_available():
        key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
    else:
        key = (-1, 0)
    ws = _WS.get(key)
    if ws is None:
        # zeroed: the first 16 KB are the split-K arrival counters of sgemm_grouped_kernel (left zero by every launch)
        ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda" if key[0] >= 0 else "cpu")
        _WS[key] = ws
    return ws


# ---- grouped GEMM launches: inside ``with gemm_group():`` the head's GEMMs (and bias-gradient column sums) are
# queued and issued as ONE ``vlnce_sgemm_grouped`` launch when the outermost group closes.  Only problems that do
# not depend on each other may share a group (the launch plans below mark those regions by hand).
_PENDING = None  # list of problem tuples while a group is open
# VLNCE_GEMM_GROUPS=1 turns the grouping on.  It is OFF by default because it was measured, not because it is
# unfinished: alone the grouped head is 59 C-ABI calls instead of 148 and ~0.3 ms shorter when its kernels run one
# after the other, but in the pipelined training step the conv trunks of the next batch are the critical path
# (tools/timeline_probe.py) and the step came out 0.03-0.05 ms (1-2 %) slower in three A/B runs on the B200 — bigger
# head launches displace more conv CTAs at once.  Both settings are covered by the GPU tests.
GROUPING = __import__("os").environ.get("VLNCE_GEMM_GROUPS", "0") != "0"
# VLNCE_DEFER_LEAVES=0: weight / bias gradients run on a side stream next to the dX chain (one launch each) instead of
# being collected into one grouped launch at the end of the backward
DEFER_LEAVES = __import__("os").environ.get("VLNCE_DEFER_LEAVES", "1") != "0"


def _flush(problems):
    if not problems:
        return
    import ctypes
    import struct

    blob = b"".join(struct.pack("<12qdq2q", a or 0, lda_m, lda_k, b or 0, ldb_k, ldb_n, c, ldc, bias or 0, m, n, k,
                                float(beta), int(act), 0, 0)
                    for (a, lda_m, lda_k, b, ldb_k, ldb_n, c, ldc, bias, m, n, k, beta, act) in problems)
    buf = ctypes.create_string_buffer(blob, len(blob))
    ws = gemm_workspace()
    C.vlnce_sgemm_grouped(ctypes.addressof(buf), len(problems), ws.data_ptr(), ws.numel(), lib.stream())


class gemm_group:
    def __enter__(self):
        global _PENDING
        self.outer = GROUPING and _PENDING is None
        if self.outer:
            _PENDING = []
        return self

    def __exit__(self, et, ev, tb):
        global _PENDING
        if self.outer:
            problems, _PENDING = _PENDING, None
            if et is None:
                _flush(problems)
        return False


def colsum(a, lda, out, m, n, acc=0):
    """out[n] (+)= sum_m a[m*lda + n] (bias gradients): a column-sum problem of the open group, else one launch."""
    if _PENDING is not None:
        _PENDING.append((a, 1, lda, None, 0, 0, out, 1, None, n, 1, m, 1.0 if acc else 0.0, 0))
        return
    ws = gemm_workspace()
    C.vlnce_colsum(a, lda, out, m, n, acc, ws.data_ptr(), ws.numel(), lib.stream())


class Branch:
    """Fork/join of an independent branch of the step onto a side stream (captured as parallel
    graph branches under CUDA-graph capture).  No-op without CUDA (CPU dry-run walks).
    ``priority``: CUDA stream priority of the side stream (-1 = high); branches opened inside a
    high-priority branch inherit it, so a whole sub-graph (the head update next to the conv trunks of
    the following batch) can be preferred by the block scheduler."""

    _streams = {}
    _inherit = 0  # priority of the innermost open branch that asked for one

    def __init__(self, name: str, enabled=None, priority=None):
        self.enabled = (OVERLAP if enabled is None else enabled) and torch.cuda.is_available()
        self.name = name
        self.ctx = None
        self.priority = Branch._inherit if priority is None else int(priority)
        if self.enabled:
            key = (torch.cuda.current_device(), name, self.priority)
            if key not in Branch._streams:
                Branch._streams[key] = torch.cuda.Stream(priority=self.priority)
            self.stream = Branch._streams[key]
            self.done = torch.cuda.Event()

    def __enter__(self):
        self._saved_inherit = Branch._inherit
        Branch._inherit = self.priority
        if self.enabled:
            self.stream.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.stream)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        Branch._inherit = self._saved_inherit
        if self.enabled:
            self.done.record(self.stream)
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.enabled:
            torch.cuda.current_stream().wait_event(self.done)


def _gemm(a, lda_m, lda_k, b, ldb_k, ldb_n, c, ldc, bias, m, n, k, beta=0.0, act=0):
    if _PENDING is not None:
        _PENDING.append((a, lda_m, lda_k, b, ldb_k, ldb_n, c, ldc, bias, m, n, k, beta, act))
        return
    ws = gemm_workspace()
    C.vlnce_sgemm(a, lda_m, lda_k, b, ldb_k, ldb_n, c, ldc, bias, m, n, k, float(beta), int(act),
                  ws.data_ptr(), ws.numel(), lib.stream())


def linear(x, ldx, w, out, ldo, bias, rows, n_out, n_in, act=0, beta=0.0, ldw=None):
    """out[rows, n_out] = act(x[rows, n_in] @ w[n_out, n_in]^T + bias) — pointers, not tensors."""
    ldw = n_in if ldw is None else ldw
    _gemm(x, ldx, 1, w, 1, ldw, out, ldo, bias, rows, n_out, n_in, beta, act)


def linear_dx(dy, lddy, w, dx, lddx, rows, n_out, n_in, beta=0.0, ldw=None):
    """dx[rows, n_in] (+)= dy[rows, n_out] @ w[n_out, n_in]"""
    ldw = n_in if ldw is None else ldw
    _gemm(dy, lddy, 1, w, ldw, 1, dx, lddx, None, rows, n_in, n_out, beta, 0)


def linear_dw(dy, lddy, x, ldx, dw, lddw, rows, n_out, n_in, beta=0.0):
    """dw[n_out, n_in] (+)= dy[rows, n_out]^T @ x[rows, n_in]"""
    _gemm(dy, 1, lddy, x, ldx, 1, dw, lddw, None, n_out, n_in, rows, beta, 0)


import os as _os

# Large head GEMMs on the tensor cores (TF32 products, fp32 accumulation; the persistent TMA tile's 1x1 mode as a GEMM):
# the instruction-feature input projection and its weight gradient at RxR size (512 tokens x 768 features: 2 x 12.9
# GFLOP each way per 32 rows — the fp32 CUDA-core GEMM made the C5 step head-bound).  Used by the reduced-precision modes
# (precision "bf16" / "tf32", 1e-2 class) for problems with at least this many rows; precision "fp32" keeps CUDA cores.
HEAD_TC_MIN_ROWS = int(_os.environ.get("VLNCE_HEAD_TC_MIN_ROWS", "2048"))
_ONES: Dict[tuple, torch.Tensor] = {}


def _ones(n: int, device) -> torch.Tensor:
    key = (n, str(device))
    t = _ONES.get(key)
    if t is None:
        t = _ONES[key] = torch.ones(n, dtype=torch.float32, device=device)
    return t


def tc_gemm_ok(precision: str, rows: int, n_out: int, n_in: int) -> bool:
    return (precision != "fp32" and HEAD_TC_MIN_ROWS > 0 and rows >= HEAD_TC_MIN_ROWS and n_in % 32 == 0
            and n_out % 64 == 0)


def linear_tc(x_r, w_r, out, bias, rows: int, n_out: int, n_in: int, device):
    """out[rows, n_out] = x_r[rows, n_in] @ w_r[n_out, n_in]^T + bias on tcgen05 (kind::tf32); all matrices contiguous,
    operands already rounded to TF32 (``vlnce_round_tf32`` / ``vlnce_transpose_f32``).  bias None = 0."""
    shift = bias if bias is not None else _zeros(n_out, device).data_ptr()
    C.vlnce_conv_tma(x_r, w_r, lib.F32, _ones(n_out, device).data_ptr(), shift, None, out, lib.F32, 1, 1, rows, n_in,
                     n_out, 1, 2, 128 if n_out % 128 == 0 else 64, lib.stream())


_ZEROS: Dict[tuple, torch.Tensor] = {}


def _zeros(n: int, device) -> torch.Tensor:
    key = (n, str(device))
    t = _ZEROS.get(key)
    if t is None:
        t = _ZEROS[key] = torch.zeros(n, dtype=torch.float32, device=device)
    return t


OVERLAP = _os.environ.get("VLNCE_OVERLAP", "1") != "0"  # multi-stream branches on/off


def set_overlap(on: bool):
    """Toggle the multi-stream branches at run time (profiling passes want serial launches)."""
    global OVERLAP
    OVERLAP = bool(on)
F4 = 4  # sizeof(float)
# single GRU steps of the state encoders as one fused launch (vlnce_gru_step_fwd); 0 = mask / two GEMMs / pointwise
FUSED_GRU_STEP = _os.environ.get("VLNCE_FUSED_GRU", "1") != "0"


def off(t: torch.Tensor, elems: int) -> int:
    """device pointer of ``t`` advanced by ``elems`` fp32 elements"""
    return t.data_ptr() + F4 * elems


class HeadEngine:
    """Shared pieces of the policy launch plans: workspace, GRU state encoder, action head."""

    def __init__(self, net, precision: str = "bf16"):
        self.net = net
        self.precision = precision
        self.pool: Optional[BufferPool] = None
        self.ctx: Dict[str, object] = {}
        self.H = net.model_config.STATE_ENCODER.hidden_size

    def buf(self, name, shape, dtype=torch.float32):
        return self.pool.get(name, shape, dtype)

    def ensure_pool(self, dev):
        if self.pool is None or self.pool.device != dev:
            self.pool = BufferPool(dev)
        return self.pool

    def _prep_rows(self, rnn_states, prev_actions, masks):
        dev = rnn_states.device
        lib.require_cuda(rnn_states, "rnn_states")
        self.ensure_pool(dev)
        N = rnn_states.shape[0]
        R = prev_actions.shape[0]
        assert R % N == 0, "rows must be T*N (time-major)"
        rnn_states = rnn_states.contiguous().float()
        masks_f = self.buf("masks_f", (R,))
        masks_f.copy_(masks.reshape(-1))  # dtype cast only (uint8/bool/float -> f32)
        prev = prev_actions.reshape(-1)
        if prev.dtype != torch.int64:
            prev = prev.long()
        return rnn_states, masks_f, prev.contiguous(), R, N, R // N

    def _encode_instruction(self, observations, R, save):
        """instruction_encoder.py:63-94: lengths, embedding, input projection, persistent LSTM / GRU.
        Returns (per-token outputs (R,L,dirs*Hs), final states (dirs,R,Hs), lengths, L)."""
        ie = self.net.instruction_encoder
        st = lib.stream()
        Hs, dirs, E = ie.config.hidden_size, (2 if ie.config.bidirectional else 1), ie.config.embedding_size
        lengths = self.buf("lengths", (R,), torch.int32)
        if ie.config.sensor_uuid == "instruction":
            ids = observations["instruction"]
            lib.require_cuda(ids, "instruction")
            ids = ids.long().contiguous()
            L = ids.shape[1]
            C.vlnce_token_lengths(ptr(ids), ptr(lengths), R, L, st)
            emb = self.buf("instr_emb", (R * L, E))
            C.vlnce_embedding_fwd(ptr(ids), ptr(ie.embedding_layer.weight), ptr(emb), R * L, E, E,
                                  ie.embedding_layer.num_embeddings, st)
            self.ctx["ids"] = ids
        else:
            emb = observations["rxr_instruction"]
            lib.require_cuda(emb, "rxr_instruction")
            emb = emb.contiguous().float()
            L = emb.shape[1]
            C.vlnce_feature_lengths(ptr(emb), ptr(lengths), R, L, E, st)
            self.ctx["rxr"] = emb
        rnn = ie.encoder_rnn
        gates = 3 if isinstance(rnn, torch.nn.GRU) else 4  # r,z,n | i,f,g,o
        gx = self.buf("lstm_gx", (dirs, R, L, gates * Hs))
        wih = [rnn.weight_ih_l0] + ([rnn.weight_ih_l0_reverse] if dirs == 2 else [])
        bih = [rnn.bias_ih_l0] + ([rnn.bias_ih_l0_reverse] if dirs == 2 else [])
        G = gates * Hs
        self.ctx["instr_tc"] = tc_gemm_ok(self.precision, R * L, G, E)
        if self.ctx["instr_tc"]:
            # tensor-core input projection: operands rounded to TF32 once (the rounded features are kept for the
            # weight gradient), one persistent TMA-tile launch per direction
            emb_r = self.buf("instr_emb_tf32", (R * L, E))
            C.vlnce_round_tf32(ptr(emb), ptr(emb_r), R * L * E, st)
            w_r = self.buf("instr_wih_tf32", (dirs, G, E))
            for d in range(dirs):
                C.vlnce_round_tf32(ptr(wih[d]), off(w_r, d * G * E), G * E, st)
                linear_tc(ptr(emb_r), off(w_r, d * G * E), off(gx, d * R * L * G), ptr(bih[d]), R * L, G, E, emb.device)
        else:
            with gemm_group():  # both directions' input projections: one launch
                for d in range(dirs):
                    linear(ptr(emb), E, ptr(wih[d]), off(gx, d * R * L * G), G, ptr(bih[d]), R * L, G, E)
        instr = self.buf("instr_out", (R, L, dirs * Hs))
        hn = self.buf("lstm_hn", (dirs, R, Hs))
        whh_b = ptr(rnn.weight_hh_l0_reverse) if dirs == 2 else None
        bhh_b = ptr(rnn.bias_hh_l0_reverse) if dirs == 2 else None
        if gates == 3:  # INSTRUCTION_ENCODER.rnn_type = "GRU": final_state is h_n itself (instruction_encoder.py:86-87)
            sv = self.buf("gru_seq_save", (dirs, R, L, 4 * Hs)) if save else None
            C.vlnce_gru_seq_fwd(ptr(gx), ptr(rnn.weight_hh_l0), whh_b, ptr(rnn.bias_hh_l0), bhh_b, ptr(lengths),
                                ptr(instr), ptr(hn), ptr(sv), R, L, Hs, dirs, st)
            return instr, hn, lengths, L
        sg = self.buf("lstm_gates", (dirs, R, L, 4 * Hs)) if save else None
        sc = self.buf("lstm_c", (dirs, R, L, Hs)) if save else None
        C.vlnce_lstm_seq_fwd(ptr(gx), ptr(rnn.weight_hh_l0), whh_b, ptr(rnn.bias_hh_l0), bhh_b, ptr(lengths),
                             ptr(instr), ptr(hn), ptr(sg), ptr(sc), R, L, Hs, dirs, st)
        return instr, hn, lengths, L

    # ------------------------------------------------------------------ state-encoder helper
    @staticmethod
    def slots(rnn) -> int:
        """rnn_states slots one state encoder owns: (h) for a GRU, (h, c) for an LSTM
        (habitat_baselines v0.1.7 ``RNNStateEncoder.num_recurrent_layers``)."""
        return 2 if isinstance(rnn, torch.nn.LSTM) else 1

    def _gru_forward(self, tag, rnn, x, ldx, in_size, h0, ldh0, masks_f, out, ldo, h_final, ldhf, T, N,
                     save):
        """habitat_baselines v0.1.7 RNNStateEncoder forward (single step when T == 1, else a
        T-step sequence with hidden reset where masks == 0); writes all T*N outputs into
        ``out`` and the final hidden state into ``h_final``.  ``h0`` / ``h_final`` point at the encoder's
        first slot of the (N, slots, H) state rows; an LSTM's cell state is the slot after it."""
        if isinstance(rnn, torch.nn.LSTM):
            return self._lstm_forward(tag, rnn, x, ldx, in_size, h0, ldh0, masks_f, out, ldo, h_final, ldhf, T, N,
                                      save)
        H = self.H
        R = T * N
        if T == 1 and FUSED_GRU_STEP and H % 4 == 0 and 12 * (in_size + H) * F4 <= 200 * 1024:
            # single step (act, T = 1 training): reset + both projections + gates + update in ONE launch
            hm_all = self.buf(tag + "_hm", (R, H)) if save else None
            sv = self.buf(tag + "_save", (R, 4 * H)) if save else None
            C.vlnce_gru_step_fwd(x, ldx, in_size, h0, ldh0, ptr(masks_f), ptr(rnn.weight_ih_l0), ptr(rnn.weight_hh_l0),
                                 ptr(rnn.bias_ih_l0), ptr(rnn.bias_hh_l0), ptr(hm_all), out, ldo, h_final, ldhf, ptr(sv),
                                 N, H, lib.stream())
            return
        gi = self.buf(tag + "_gi", (R, 3 * H))
        hm_all = self.buf(tag + "_hm", (R, H))
        gh = self.buf(tag + "_gh", (N, 3 * H))
        sv = self.buf(tag + "_save", (R, 4 * H)) if save else None
        st = lib.stream()
        for t in range(T):
            hm = off(hm_all, t * N * H)
            if t == 0:
                C.vlnce_mask_rows(h0, ldh0, off(masks_f, 0), hm, H, N, H, st)
                with gemm_group():  # input projection of every step + hidden projection of step 0: independent
                    linear(x, ldx, ptr(rnn.weight_ih_l0), ptr(gi), 3 * H, ptr(rnn.bias_ih_l0), R, 3 * H, in_size)
                    linear(hm, H, ptr(rnn.weight_hh_l0), ptr(gh), 3 * H, ptr(rnn.bias_hh_l0), N, 3 * H, H)
            else:
                C.vlnce_mask_rows(out + F4 * (t - 1) * N * ldo, ldo, off(masks_f, t * N), hm, H, N, H, st)
                linear(hm, H, ptr(rnn.weight_hh_l0), ptr(gh), 3 * H, ptr(rnn.bias_hh_l0), N, 3 * H, H)
            last = t == T - 1
            C.vlnce_gru_pointwise_fwd(off(gi, t * N * 3 * H), ptr(gh), hm, H, out + F4 * t * N * ldo, ldo,
                                      h_final if last else None, ldhf,
                                      off(sv, t * N * 4 * H) if save else None, N, H, st)

    def _lstm_forward(self, tag, rnn, x, ldx, in_size, h0, ldh0, masks_f, out, ldo, h_final, ldhf, T, N, save):
        """``STATE_ENCODER.rnn_type = "LSTM"``: the packed hidden state is ``cat([h, c], 0)``
        (rnn_state_encoder ``_pack_hidden``), i.e. slots (h, c) of each row; both are reset where masks == 0."""
        H = self.H
        R = T * N
        gi = self.buf(tag + "_gi", (R, 4 * H))
        linear(x, ldx, ptr(rnn.weight_ih_l0), ptr(gi), 4 * H, ptr(rnn.bias_ih_l0), R, 4 * H, in_size)
        hm_all, cm_all = self.buf(tag + "_hm", (R, H)), self.buf(tag + "_cm", (R, H))
        c_all = self.buf(tag + "_c", (R, H))
        gh = self.buf(tag + "_gh", (N, 4 * H))
        sv = self.buf(tag + "_save", (R, 5 * H)) if save else None
        st = lib.stream()
        for t in range(T):
            hm, cm = off(hm_all, t * N * H), off(cm_all, t * N * H)
            if t == 0:
                C.vlnce_mask_rows(h0, ldh0, off(masks_f, 0), hm, H, N, H, st)
                C.vlnce_mask_rows(h0 + F4 * H, ldh0, off(masks_f, 0), cm, H, N, H, st)
            else:
                C.vlnce_mask_rows(out + F4 * (t - 1) * N * ldo, ldo, off(masks_f, t * N), hm, H, N, H, st)
                C.vlnce_mask_rows(off(c_all, (t - 1) * N * H), H, off(masks_f, t * N), cm, H, N, H, st)
            linear(hm, H, ptr(rnn.weight_hh_l0), ptr(gh), 4 * H, ptr(rnn.bias_hh_l0), N, 4 * H, H)
            last = t == T - 1
            C.vlnce_lstm_pointwise_fwd(off(gi, t * N * 4 * H), ptr(gh), cm, H, out + F4 * t * N * ldo, ldo,
                                       off(c_all, t * N * H), H, h_final if last else None,
                                       h_final + F4 * H if last else None, ldhf,
                                       off(sv, t * N * 5 * H) if save else None, N, H, st)

    # ------------------------------------------------------------------ action head
    def action_logits(self, head_linear, x_out, R):
        """CategoricalNet (policy.py:19-21,37): Linear then the log-softmax Categorical applies."""
        A = head_linear.weight.shape[0]
        logits = self.buf("logits", (R, A))
        linear(ptr(x_out), self.H, ptr(head_linear.weight), ptr(logits), A, ptr(head_linear.bias), R, A, self.H)
        C.vlnce_log_softmax_rows(ptr(logits), R, A, lib.stream())
        return logits


class CMAEngine(HeadEngine):
    """Forward/backward launch plan bound to one ``CMANet`` (+ action head)."""

    def __init__(self, net, precision: str = "bf16"):
        super().__init__(net, precision)
        cfg = net.model_config
        self.H = cfg.STATE_ENCODER.hidden_size
        self.rgb_out = cfg.RGB_ENCODER.output_size
        self.depth_out = cfg.DEPTH_ENCODER.output_size
        ie = net.instruction_encoder
        self.Hs = ie.config.hidden_size
        self.dirs = 2 if ie.config.bidirectional else 1
        self.E = ie.config.embedding_size
        self.Dt = self.Hs * self.dirs  # instruction feature size
        self.Crgb = net.rgb_encoder.output_shape[0]      # 2112
        self.Cdep = net.depth_encoder.output_shape[0]    # 192
        self.P = 16
        self.state_in = self.rgb_out + self.depth_out + 32
        self.cat = self.H + self.Dt + self.rgb_out + self.depth_out + 32

    def run_backward(self, head_linear, g, d_raw, d_x, d_hat):
        """Hand-written backward of the last ``forward(save=True)`` (backward.cma_backward)."""
        from .backward import cma_backward

        return cma_backward(self, head_linear, g, d_raw, d_x, d_hat)

    # ------------------------------------------------------------------ forward
    def visual_buffers(self, R, slot: str):
        """(rgb (R,16,2112), depth (R,16,192)) fp32 feature buffers of a named slot."""
        return (self.buf(f"rgb_emb_{slot}", (R, self.P, self.Crgb)), self.buf(f"depth_emb_{slot}", (R, self.P, self.Cdep)))

    def encode_visual(self, observations, rgb_e, dep_e, save: bool = False):
        """Both frozen conv trunks (depth on a forked stream) -> channels [0, C_trunk) of the
        feature buffers.  Independent of every trainable parameter, so a training loop may run it
        for batch i+1 while the head of batch i trains (``ILTrainerStep.pipelined_step``)."""
        net = self.net
        self.ensure_pool(rgb_e.device)
        idx = getattr(self, "valid_rows", None)
        if idx is not None and self._can_skip_rows(observations):
            # teacher-forcing batches are padded to the longest trajectory (dagger_trainer.py:39-114): rows with weight
            # 0 reach no loss term and no valid row's recurrent state (they follow their trajectory's last real step),
            # so the frozen eval-mode trunks only see the real rows; the padded rows' features stay zero
            sub = dict(observations, rgb=observations["rgb"].index_select(0, idx),
                       depth=observations["depth"].index_select(0, idx))
            n = idx.numel()
            rc, dc = self.buf("rgb_emb_valid", (n, self.P, self.Crgb)), self.buf("depth_emb_valid", (n, self.P, self.Cdep))
            with Branch("depth") as br_depth:
                net.depth_encoder.features_into(sub, dc, self.Cdep, self.precision, save=False)
            net.rgb_encoder.features_into(sub, rc, self.Crgb, self.precision, save=False)
            br_depth.join()
            rgb_e.zero_()
            dep_e.zero_()
            rgb_e.index_copy_(0, idx, rc)
            dep_e.index_copy_(0, idx, dc)
            return
        with Branch("depth") as br_depth:
            net.depth_encoder.features_into(observations, dep_e, self.Cdep, self.precision, save=save)
        net.rgb_encoder.features_into(observations, rgb_e, self.Crgb, self.precision, save=save)
        br_depth.join()

    def _can_skip_rows(self, observations) -> bool:
        """padded rows may bypass the trunks only if no image's features depend on another image (eval-mode BatchNorm;
        GroupNorm is per image), the trunks are frozen, raw frames are the input and nobody hooks the trunk outputs."""
        net = self.net
        rgb_enc, dep_enc = net.rgb_encoder, net.depth_encoder
        if "rgb" not in observations or "depth" not in observations:
            return False
        if getattr(rgb_enc, "trainable", False) or getattr(dep_enc, "trainable", False):
            return False
        cnn = getattr(rgb_enc, "cnn", None)
        if cnn is None or getattr(cnn, "training", True):
            return False
        for m in (cnn, getattr(dep_enc, "visual_encoder", None)):
            if m is not None and len(getattr(m, "_forward_hooks", {})) > 0:
                return False
        return True

    def forward(self, observations, rnn_states, prev_actions, masks, save: bool, visual=None):
        net = self.net
        rnn_states, masks_f, prev, R, N, T = self._prep_rows(rnn_states, prev_actions, masks)
        st = lib.stream()
        H, P = self.H, self.P

        # ---- three independent producers run concurrently: instruction encoder and depth trunk
        # on side streams, RGB trunk on the caller's stream (joined where first consumed)
        with Branch("instr") as br_instr:
            instr, _, lengths, L = self._encode_instruction(observations, R, save)

        # ---- visual encoders + spatial embeddings (cma_policy.py:227-231)
        c_rgb = net.rgb_encoder.resnet_layer_size
        c_dep = net.depth_encoder.visual_encoder.output_shape[0]
        if visual is None:
            rgb_e = self.buf("rgb_emb", (R, P, self.Crgb))
            dep_e = self.buf("depth_emb", (R, P, self.Cdep))
            self.encode_visual(observations, rgb_e, dep_e, save=save)
        else:  # trunk features were produced ahead of time (pipelined training loop)
            rgb_e, dep_e = visual
        C.vlnce_spatial_embed_fill(ptr(net.depth_encoder.spatial_embeddings.weight), ptr(dep_e),
                                   P * self.Cdep, self.Cdep, c_dep, R, P, 64, st)
        C.vlnce_spatial_embed_fill(ptr(net.rgb_encoder.spatial_embeddings.weight), ptr(rgb_e), P * self.Crgb,
                                   self.Crgb, c_rgb, R, P, 64, st)
        cfg = net.model_config
        # modality ablations (cma_policy.py:237-242): ``embedding * 0`` — the buffers are cleared; the backward
        # leaves the gradients of everything upstream of a cleared embedding at zero (backward.cma_backward)
        if cfg.get("ablate_depth", False):
            dep_e.zero_()
        if cfg.get("ablate_rgb", False):
            rgb_e.zero_()
        self.ctx["rgb_e"], self.ctx["dep_e"] = rgb_e, dep_e

        # ---- prev-action embedding into both concat slices (cma_policy.py:233-235, Q6)
        state_in = self.buf("state_in", (R, self.state_in))
        xcat = self.buf("xcat", (R, self.cat))
        pa_w = net.prev_action_embedding.weight
        C.vlnce_prev_action_embed(ptr(prev), ptr(masks_f), ptr(pa_w), off(state_in, self.rgb_out + self.depth_out),
                                  self.state_in, off(xcat, self.cat - 32), self.cat, R, 32, pa_w.shape[0], st)
        self.ctx["prev"] = prev

        # ---- rgb_linear / depth_linear (cma_policy.py:244-245)
        rgb_mean = self.buf("rgb_mean", (R, self.Crgb))
        C.vlnce_mean_positions(ptr(rgb_e), P * self.Crgb, self.Crgb, ptr(rgb_mean), self.Crgb, R, P, self.Crgb, st)
        dep_flat = self.buf("depth_flat", (R, self.Cdep * P))
        C.vlnce_positions_to_nchw(ptr(dep_e), P * self.Cdep, self.Cdep, ptr(dep_flat), self.Cdep * P, R, P,
                                  self.Cdep, st)
        with gemm_group():  # independent: one launch
            lin = net.rgb_linear[2]
            linear(ptr(rgb_mean), self.Crgb, ptr(lin.weight), ptr(state_in), self.state_in, ptr(lin.bias), R,
                   self.rgb_out, self.Crgb, act=1)
            lin = net.depth_linear[1]
            linear(ptr(dep_flat), self.Cdep * P, ptr(lin.weight), off(state_in, self.rgb_out), self.state_in,
                   ptr(lin.bias), R, self.depth_out, self.Cdep * P, act=1)

        # ---- first GRU -> xcat[:, 0:H]  (cma_policy.py:247-256)
        s1, Lr = self.slots(net.state_encoder.rnn), net.num_recurrent_layers  # (N, Lr, H): encoder 1 | encoder 2
        h_out = self.buf("rnn_states_out", (N, Lr, H))
        self._gru_forward("gru1", net.state_encoder.rnn, ptr(state_in), self.state_in, self.state_in,
                          ptr(rnn_states), Lr * H, masks_f, ptr(xcat), self.cat, ptr(h_out), Lr * H, T, N, save)

        # ---- text attention (cma_policy.py:258-263)
        br_instr.join()
        if net.model_config.get("ablate_instruction", False):
            instr.zero_()
        if "scale" not in self.__dict__:
            self.scale = float(net._scale)  # one host sync, cached
        scale = self.scale
        hk = H // 2
        q_s = self.buf("q_state", (R, hk))
        linear(ptr(xcat), self.cat, ptr(net.state_q.weight), ptr(q_s), hk, ptr(net.state_q.bias), R, hk, H)
        q_t = self.buf("q_text", (R, self.Dt))
        linear_dx(ptr(q_s), hk, ptr(net.text_k.weight), ptr(q_t), self.Dt, R, hk, self.Dt)  # W_k^T q
        p_text = self.buf("p_text", (R, L)) if save else None
        C.vlnce_attn_pool_fwd(ptr(q_t), self.Dt, ptr(instr), L * self.Dt, self.Dt, ptr(lengths), scale,
                              ptr(p_text), off(xcat, H), self.cat, R, L, self.Dt, st)

        # ---- visual attentions (cma_policy.py:265-274)
        text_q = self.buf("text_q", (R, hk))
        linear(off(xcat, H), self.cat, ptr(net.text_q.weight), ptr(text_q), hk, ptr(net.text_q.bias), R, hk, self.Dt)
        vis = (("rgb", rgb_e, self.Crgb, net.rgb_kv, self.rgb_out, H + self.Dt),
               ("depth", dep_e, self.Cdep, net.depth_kv, self.depth_out, H + self.Dt + self.rgb_out))
        with gemm_group():  # W_k^T text_q for both modalities (W_k = rows [0, hk))
            for name, E_, Cc, kv, v_out, col in vis:
                linear_dx(ptr(text_q), hk, ptr(kv.weight), ptr(self.buf("q_" + name, (R, Cc))), Cc, R, hk, Cc)
        for name, E_, Cc, kv, v_out, col in vis:
            probs = self.buf("p_" + name, (R, P)) if save else None
            C.vlnce_attn_pool_fwd(ptr(self.buf("q_" + name, (R, Cc))), Cc, ptr(E_), P * Cc, Cc, None, scale, ptr(probs),
                                  ptr(self.buf("pooled_" + name, (R, Cc))), Cc, R, P, Cc, st)
        with gemm_group():  # W_v pooled + b_v for both modalities (W_v = rows [hk, hk+v_out))
            for name, E_, Cc, kv, v_out, col in vis:
                linear(ptr(self.buf("pooled_" + name, (R, Cc))), Cc, off(kv.weight, hk * Cc), off(xcat, col), self.cat,
                       off(kv.bias, hk), R, v_out, Cc)

        # ---- compress + second GRU (cma_policy.py:276-294)
        x2 = self.buf("x2", (R, H))
        lin = net.second_state_compress[0]
        linear(ptr(xcat), self.cat, ptr(lin.weight), ptr(x2), H, ptr(lin.bias), R, H, self.cat, act=1)
        x_out = self.buf("x_out", (R, H))
        self._gru_forward("gru2", net.second_state_encoder.rnn, ptr(x2), H, H, off(rnn_states, s1 * H), Lr * H,
                          masks_f, ptr(x_out), H, off(h_out, s1 * H), Lr * H, T, N, save)

        # ---- progress monitor (cma_policy.py:296) — tanh(Linear(x)); the loss lives in the caller
        hat = None
        if net.model_config.PROGRESS_MONITOR.use:
            hat = self.buf("progress_hat", (R, 1))
            pm = net.progress_monitor
            linear(ptr(x_out), H, ptr(pm.weight), ptr(hat), 1, ptr(pm.bias), R, 1, H, act=2)
        self.ctx.update(R=R, N=N, T=T, L=L, lengths=lengths, saved=save)
        return x_out, h_out, hat

