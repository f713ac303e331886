"""Head kernels on the B200 against plain torch fp32 (the per-op references)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from vlnce_b200 import lib
from vlnce_b200.engine import linear, linear_dw, linear_dx
from vlnce_b200.lib import C, ptr

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _r(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("m,n,k", [(64, 256, 2112), (70, 33, 50), (1, 4, 512), (130, 1536, 416)])
def test_sgemm_variants(m, n, k):
    x, w, b = _r(m, k), _r(n, k, scale=0.1), _r(n)
    out = torch.full((m, n + 3), float("nan"), device=DEV)
    linear(ptr(x), k, ptr(w), ptr(out), n + 3, ptr(b), m, n, k, act=1)
    torch.testing.assert_close(out[:, :n], torch.relu(x @ w.t() + b), rtol=1e-4, atol=1e-4)
    dy = _r(m, n, seed=3)
    dx = torch.empty(m, k, device=DEV)
    linear_dx(ptr(dy), n, ptr(w), ptr(dx), k, m, n, k)
    torch.testing.assert_close(dx, dy @ w, rtol=1e-4, atol=1e-4)
    dw = torch.ones(n, k, device=DEV)
    linear_dw(ptr(dy), n, ptr(x), k, ptr(dw), k, m, n, k, beta=1.0)
    torch.testing.assert_close(dw, 1 + dy.t() @ x, rtol=1e-4, atol=2e-4)
    cs = torch.empty(n, device=DEV)
    from vlnce_b200.engine import gemm_workspace
    ws = gemm_workspace()
    C.vlnce_colsum(ptr(dy), n, ptr(cs), m, n, 0, ws.data_ptr(), ws.numel(), lib.stream())
    big = _r(5120, 512, seed=4)
    cs2 = torch.empty(512, device=DEV)
    C.vlnce_colsum(ptr(big), 512, ptr(cs2), 5120, 512, 0, ws.data_ptr(), ws.numel(), lib.stream())
    torch.testing.assert_close(cs2, big.sum(0), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(cs, dy.sum(0), rtol=1e-4, atol=1e-4)


def test_sgemm_grouped_matches_torch_and_is_deterministic():
    """One grouped launch = the same results as the problems one by one: mixed shapes (split-K and not), a fused
    accumulate (beta = 1), activations, and bias-gradient column sums riding along; bit-identical when repeated
    (the in-launch split-K reduction sums the partials in a fixed order whatever CTA arrives last)."""
    from vlnce_b200.engine import colsum, gemm_group

    x1, w1, b1 = _r(64, 2112), _r(256, 2112, scale=0.1), _r(256)
    x2, w2, b2 = _r(64, 3072, seed=1), _r(128, 3072, seed=1, scale=0.1), _r(128, seed=1)
    dy, xs = _r(5120, 512, seed=2, scale=0.1), _r(5120, 50, seed=2)
    d3 = _r(64, 7, seed=3)

    def run():
        o1 = torch.full((64, 256 + 128), float("nan"), device=DEV)
        dw = torch.ones(512, 50, device=DEV)
        db = torch.full((512,), float("nan"), device=DEV)
        db3 = torch.ones(7, device=DEV)
        with gemm_group():
            linear(ptr(x1), 2112, ptr(w1), ptr(o1), 384, ptr(b1), 64, 256, 2112, act=1)
            linear(ptr(x2), 3072, ptr(w2), ptr(o1) + 4 * 256, 384, ptr(b2), 64, 128, 3072, act=2)
            linear_dw(ptr(dy), 512, ptr(xs), 50, ptr(dw), 50, 5120, 512, 50, beta=1.0)
            colsum(ptr(dy), 512, ptr(db), 5120, 512)
            colsum(ptr(d3), 7, ptr(db3), 64, 7, acc=1)
        return o1, dw, db, db3

    o1, dw, db, db3 = run()
    torch.testing.assert_close(o1[:, :256], torch.relu(x1 @ w1.t() + b1), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(o1[:, 256:], torch.tanh(x2 @ w2.t() + b2), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(dw, 1 + dy.t() @ xs, rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db, dy.sum(0), rtol=1e-4, atol=1e-3)
    torch.testing.assert_close(db3, 1 + d3.sum(0), rtol=1e-4, atol=1e-4)
    for _ in range(3):
        again = run()
        assert all(torch.equal(a, b) for a, b in zip(again, (o1, dw, db, db3)))
    # more problems than one launch holds (24): issued as consecutive launches
    outs = [torch.empty(8, 8, device=DEV) for _ in range(30)]
    a = _r(8, 40, seed=5)
    with gemm_group():
        for i, o in enumerate(outs):
            linear(ptr(a), 40, ptr(w1) + 4 * 2112 * i, ptr(o), 8, None, 8, 8, 40, ldw=2112)
    for i, o in enumerate(outs):
        torch.testing.assert_close(o, a @ w1[i:i + 8, :40].t(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dirs,R,nr", [(1, 5, 1), (2, 5, 1), (2, 21, 4), (1, 64, 4), (2, 21, 1), (2, 21, 2)])
def test_lstm_matches_packed_torch(dirs, R, nr):
    """Persistent (bi)LSTM forward AND backward-through-time against torch.nn.LSTM on a packed ragged batch (CPU fp32
    autograd), with one row per CTA (the default) and with four rows per CTA (``vlnce_set_lstm_rows(4)``; R = 21 has a
    ragged tail group) — lengths differ inside every group, including a length-1 row and full-length rows."""
    C.vlnce_set_lstm_rows(nr)
    try:
        _lstm_case(dirs, R)
    finally:
        C.vlnce_set_lstm_rows(0)


def _lstm_case(dirs, R):
    L, E, H = 11, 50, 128
    torch.manual_seed(0)
    rnn = nn.LSTM(E, H, bidirectional=dirs == 2)
    x = _r(R, L, E).cpu()
    lens = torch.tensor(([11, 1, 7, 4, 9, 2, 11, 3] * 8)[:R])
    x = x * (torch.arange(L).view(1, L, 1) < lens.view(R, 1, 1))
    packed = nn.utils.rnn.pack_padded_sequence(x, lens, batch_first=True, enforce_sorted=False)
    o, (hn, _) = rnn(packed)  # CPU fp32 reference (cuDNN RNNs default to TF32)
    want = nn.utils.rnn.pad_packed_sequence(o, batch_first=True, total_length=L)[0]
    d_out = torch.randn(R, L, dirs * H, generator=torch.Generator().manual_seed(3)) * (
        torch.arange(L).view(1, L, 1) < lens.view(R, 1, 1))
    (want * d_out).sum().backward()
    grads = {k: p.grad.clone().to(DEV) for k, p in rnn.named_parameters()}
    want, hn, x, rnn, d_out = want.detach().to(DEV), hn.detach().to(DEV), x.to(DEV), rnn.to(DEV), d_out.to(DEV)
    lengths = lens.to(DEV).int()
    gx = torch.empty(dirs, R, L, 4 * H, device=DEV)
    sfx = ["", "_reverse"][:dirs]
    for d, sf in enumerate(sfx):
        linear(ptr(x), E, ptr(getattr(rnn, "weight_ih_l0" + sf)), gx[d].data_ptr(), 4 * H,
               ptr(getattr(rnn, "bias_ih_l0" + sf)), R * L, 4 * H, E)
    out = torch.full((R, L, dirs * H), float("nan"), device=DEV)
    hn_got = torch.empty(dirs, R, H, device=DEV)
    gates = torch.zeros(dirs, R, L, 4 * H, device=DEV)
    cs = torch.zeros(dirs, R, L, H, device=DEV)
    whh_b = ptr(rnn.weight_hh_l0_reverse) if dirs == 2 else None
    C.vlnce_lstm_seq_fwd(ptr(gx), ptr(rnn.weight_hh_l0), whh_b,
                         ptr(rnn.bias_hh_l0), ptr(rnn.bias_hh_l0_reverse) if dirs == 2 else None, ptr(lengths),
                         ptr(out), ptr(hn_got), ptr(gates), ptr(cs), R, L, H, dirs, lib.stream())
    torch.testing.assert_close(out, want, rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(hn_got, hn, rtol=1e-4, atol=2e-5)
    # backward: pre-activation gate gradients and the hidden states fed to each step -> every parameter gradient
    dgx = torch.full((dirs, R, L, 4 * H), float("nan"), device=DEV)
    hprev = torch.full((dirs, R, L, H), float("nan"), device=DEV)
    C.vlnce_lstm_seq_bwd(ptr(d_out), ptr(gates), ptr(cs), ptr(out), ptr(rnn.weight_hh_l0), whh_b, ptr(lengths),
                         ptr(dgx), ptr(hprev), R, L, H, dirs, lib.stream())
    for d, sf in enumerate(sfx):
        dg = dgx[d].reshape(R * L, 4 * H)
        torch.testing.assert_close(dg.t() @ x.reshape(R * L, E), grads["weight_ih_l0" + sf], rtol=1e-3, atol=2e-4)
        torch.testing.assert_close(dg.t() @ hprev[d].reshape(R * L, H), grads["weight_hh_l0" + sf], rtol=1e-3, atol=2e-4)
        torch.testing.assert_close(dg.sum(0), grads["bias_ih_l0" + sf], rtol=1e-3, atol=2e-4)


def test_gru_cell_matches_torch():
    N, I, H = 7, 416, 512
    torch.manual_seed(0)
    rnn = nn.GRU(I, H)
    x, h = _r(N, I), _r(N, H, seed=5)
    with torch.no_grad():  # CPU fp32 reference
        _, want = rnn(x.cpu().unsqueeze(0), h.cpu().unsqueeze(0))
    want, rnn = want.to(DEV), rnn.to(DEV)
    gi, gh = torch.empty(N, 3 * H, device=DEV), torch.empty(N, 3 * H, device=DEV)
    linear(ptr(x), I, ptr(rnn.weight_ih_l0), ptr(gi), 3 * H, ptr(rnn.bias_ih_l0), N, 3 * H, I)
    linear(ptr(h), H, ptr(rnn.weight_hh_l0), ptr(gh), 3 * H, ptr(rnn.bias_hh_l0), N, 3 * H, H)
    out = torch.empty(N, H, device=DEV)
    C.vlnce_gru_pointwise_fwd(ptr(gi), ptr(gh), ptr(h), H, ptr(out), H, None, 0, None, N, H, lib.stream())
    torch.testing.assert_close(out, want[0], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("m,n,k", [(4096, 512, 768), (1000, 192, 96), (1024, 768, 2048)])
def test_tensor_core_head_gemm_matches_fp64(m, n, k):
    """engine.linear_tc: the persistent TMA tile's 1x1 mode as a TF32 GEMM (operands rounded to nearest TF32 by
    vlnce_round_tf32 / vlnce_transpose_f32, fp32 accumulation, plain fp32 result).  Error model: two 2^-11 operand
    roundings per product, random signs -> ~1e-3 of the result's scale at k ~ 1000."""
    from vlnce_b200.engine import linear_tc
    x, w, b = _r(m, k), _r(n, k, scale=k ** -0.5), _r(n)
    xr, wr = torch.empty_like(x), torch.empty_like(w)
    C.vlnce_round_tf32(ptr(x), ptr(xr), x.numel(), lib.stream())
    C.vlnce_round_tf32(ptr(w), ptr(wr), w.numel(), lib.stream())
    assert ((xr.view(torch.int32) & 0x1FFF) == 0).all() and (xr - x).abs().max() <= x.abs().max() * 2.0 ** -11
    out = torch.full((m, n), float("nan"), device=DEV)
    linear_tc(ptr(xr), ptr(wr), ptr(out), ptr(b), m, n, k, x.device)
    want = (x.double() @ w.double().t() + b.double()).float()
    err = (out - want).abs().max().item() / want.abs().max().item()
    print(f"tf32 gemm {m}x{n}x{k}: max err / max |y| = {err:.2e}")
    assert torch.isfinite(out).all() and err < 3e-3
    exact = (xr.double() @ wr.double().t() + b.double()).float()        # same rounded operands: only fp32 accumulation left
    assert (out - exact).abs().max().item() <= 2e-5 * want.abs().max().item()
    # the weight-gradient form: transposed (token-minor) operands, no bias
    dy = _r(m, n, seed=7)
    dyt, xt = torch.empty(n, m, device=DEV), torch.empty(k, m, device=DEV)
    C.vlnce_transpose_f32(ptr(dy), n, ptr(dyt), m, m, n, 1, lib.stream())
    C.vlnce_transpose_f32(ptr(xr), k, ptr(xt), m, m, k, 0, lib.stream())
    assert torch.equal(xt, xr.t().contiguous())
    if m % 32 == 0 and k % 64 == 0:
        dw = torch.full((n, k), float("nan"), device=DEV)
        linear_tc(ptr(dyt), ptr(xt), ptr(dw), None, n, k, m, x.device)
        want_dw = (dy.double().t() @ x.double()).float()
        err = (dw - want_dw).abs().max().item() / want_dw.abs().max().item()
        print(f"tf32 weight-gradient gemm: max err / max |dW| = {err:.2e}")
        assert err < 3e-3


@pytest.mark.parametrize("N,I,H", [(64, 416, 512), (7, 1184, 512), (1, 33, 64), (130, 50, 128)])
def test_fused_gru_step_matches_torch_and_unfused(N, I, H):
    """vlnce_gru_step_fwd (reset + both projections + gates + update in one launch) against torch.nn.GRU on the CPU and
    against the unfused kernel sequence it replaces, including what the backward reads (masked hidden, saved gates)."""
    torch.manual_seed(1)
    rnn = nn.GRU(I, H)
    x, h = _r(N, I), _r(N, 2 * H, seed=5)              # h0 rows strided like the (N, slots, H) state tensor
    mask = (torch.arange(N) % 3 != 0).float().to(DEV)    # every third row restarts an episode
    with torch.no_grad():
        _, want = rnn(x.cpu().unsqueeze(0), (h[:, :H] * mask[:, None]).cpu().unsqueeze(0))
    want, rnn = want[0].to(DEV), rnn.to(DEV)
    out = torch.full((N, H + 5), float("nan"), device=DEV)
    out2 = torch.full((N, 2 * H), float("nan"), device=DEV)
    hm, save = torch.empty(N, H, device=DEV), torch.empty(N, 4 * H, device=DEV)
    C.vlnce_gru_step_fwd(ptr(x), I, I, ptr(h), 2 * H, ptr(mask), ptr(rnn.weight_ih_l0), ptr(rnn.weight_hh_l0),
                         ptr(rnn.bias_ih_l0), ptr(rnn.bias_hh_l0), ptr(hm), ptr(out), H + 5, ptr(out2), 2 * H, ptr(save),
                         N, H, lib.stream())
    torch.testing.assert_close(out[:, :H], want, rtol=1e-4, atol=2e-5)
    assert torch.equal(out2[:, :H], out[:, :H]) and torch.isnan(out2[:, H:]).all() and torch.isnan(out[:, H:]).all()
    assert torch.equal(hm, h[:, :H] * mask[:, None])
    # the unfused sequence
    hm_u = torch.empty(N, H, device=DEV)
    C.vlnce_mask_rows(ptr(h), 2 * H, ptr(mask), ptr(hm_u), H, N, H, lib.stream())
    gi, gh = torch.empty(N, 3 * H, device=DEV), torch.empty(N, 3 * H, device=DEV)
    linear(ptr(x), I, ptr(rnn.weight_ih_l0), ptr(gi), 3 * H, ptr(rnn.bias_ih_l0), N, 3 * H, I)
    linear(ptr(hm_u), H, ptr(rnn.weight_hh_l0), ptr(gh), 3 * H, ptr(rnn.bias_hh_l0), N, 3 * H, H)
    out_u, save_u = torch.empty(N, H, device=DEV), torch.empty(N, 4 * H, device=DEV)
    C.vlnce_gru_pointwise_fwd(ptr(gi), ptr(gh), ptr(hm_u), H, ptr(out_u), H, None, 0, ptr(save_u), N, H, lib.stream())
    torch.testing.assert_close(out[:, :H], out_u, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(save, save_u, rtol=1e-4, atol=1e-5)
    # inference form: no saved tensors
    out3 = torch.empty(N, H, device=DEV)
    C.vlnce_gru_step_fwd(ptr(x), I, I, ptr(h), 2 * H, ptr(mask), ptr(rnn.weight_ih_l0), ptr(rnn.weight_hh_l0),
                         ptr(rnn.bias_ih_l0), ptr(rnn.bias_hh_l0), None, ptr(out3), H, None, 0, None, N, H, lib.stream())
    assert torch.equal(out3, out[:, :H])


@pytest.mark.parametrize("P,D,masked", [(80, 256, True), (16, 2112, False), (16, 192, False)])
def test_attention_pool_matches_reference_attn(P, D, masked):
    """cma_policy.py:207-217 with keys/values from a Conv1d, against the re-associated kernel."""
    R, hk, dv = 6, 256, 128
    E = _r(R, P, D, scale=0.5)
    lens = torch.randint(1, P + 1, (R,), generator=torch.Generator().manual_seed(1)).to(DEV)
    if masked:
        E = E * (torch.arange(P, device=DEV).view(1, P, 1) < lens.view(R, 1, 1))
    wk, bk, wv, bv = _r(hk, D, scale=D ** -0.5), _r(hk), _r(dv, D, scale=D ** -0.5), _r(dv)
    q = _r(R, hk, seed=9)
    scale = 1.0 / 16
    En = E.permute(0, 2, 1)  # (R, D, P) reference layout
    k = torch.einsum("od,rdp->rop", wk, En) + bk.view(1, -1, 1)
    v = (torch.einsum("od,rdp->rop", wv, En) + bv.view(1, -1, 1)) if not masked else En
    logits = torch.einsum("nc,nci->ni", q, k)
    if masked:
        logits = logits - (En == 0).all(1).float() * 1e8
    want = torch.einsum("ni,nci->nc", F.softmax(logits * scale, 1), v)
    qp = q @ wk
    pooled = torch.empty(R, D, device=DEV)
    probs = torch.empty(R, P, device=DEV)
    C.vlnce_attn_pool_fwd(ptr(qp), D, ptr(E), P * D, D, ptr(lens.int()) if masked else None, scale, ptr(probs),
                          ptr(pooled), D, R, P, D, lib.stream())
    got = pooled if masked else pooled @ wv.t() + bv
    torch.testing.assert_close(got, want, rtol=2e-4, atol=2e-5)
    torch.testing.assert_close(probs.sum(1), torch.ones(R, device=DEV), rtol=1e-5, atol=1e-5)


def test_small_layout_ops():
    R, P, Cc = 3, 16, 192
    E = _r(R, P, Cc)
    out = torch.empty(R, Cc * P, device=DEV)
    C.vlnce_positions_to_nchw(ptr(E), P * Cc, Cc, ptr(out), Cc * P, R, P, Cc, lib.stream())
    torch.testing.assert_close(out, E.permute(0, 2, 1).reshape(R, -1))
    f = _r(R, 128, 4, 4)
    dst = torch.zeros(R, P, Cc, device=DEV)
    C.vlnce_nchw_to_positions(ptr(f), ptr(dst), Cc, R, 128, 16, lib.stream())
    torch.testing.assert_close(dst[:, :, :128], f.flatten(2).permute(0, 2, 1))
    table = _r(16, 64)
    C.vlnce_spatial_embed_fill(ptr(table), ptr(dst), P * Cc, Cc, 128, R, P, 64, lib.stream())
    want = table.view(1, 64, 4, 4).expand(R, 64, 4, 4).flatten(2).permute(0, 2, 1)
    torch.testing.assert_close(dst[:, :, 128:], want)
    m = torch.empty(R, Cc, device=DEV)
    C.vlnce_mean_positions(ptr(dst), P * Cc, Cc, ptr(m), Cc, R, P, Cc, lib.stream())
    torch.testing.assert_close(m, dst.mean(1), rtol=1e-5, atol=1e-6)
    x = _r(4, 9, 9, 40)
    y = torch.empty(4, 16, 40, device=DEV)
    C.vlnce_adaptive_avgpool(ptr(x), lib.F32, ptr(y), 40, 4, 9, 9, 40, 4, lib.stream())
    want = F.adaptive_avg_pool2d(x.permute(0, 3, 1, 2), (4, 4)).flatten(2).permute(0, 2, 1)
    torch.testing.assert_close(y, want, rtol=1e-5, atol=1e-6)
    for dt, code in ((torch.bfloat16, lib.BF16), (torch.float16, lib.F16)):   # 8-channel vector path, strided output
        xb = _r(3, 8, 8, 64, seed=9).to(dt)
        yb = torch.full((3, 16, 72), float("nan"), device=DEV)
        C.vlnce_adaptive_avgpool(ptr(xb), code, ptr(yb), 72, 3, 8, 8, 64, 4, lib.stream())
        want = F.adaptive_avg_pool2d(xb.float().permute(0, 3, 1, 2), (4, 4)).flatten(2).permute(0, 2, 1)
        torch.testing.assert_close(yb[:, :, :64], want, rtol=1e-5, atol=1e-6)
        assert torch.isnan(yb[:, :, 64:]).all()
    ids = torch.tensor([[3, 0, 5, 0], [0, 0, 0, 0], [1, 2, 3, 4]], device=DEV)
    lens = torch.empty(3, dtype=torch.int32, device=DEV)
    C.vlnce_token_lengths(ptr(ids), ptr(lens), 3, 4, lib.stream())
    assert lens.tolist() == [2, 0, 4]


def test_inflection_weights_and_device_collate():
    """the inflection-weight kernel on the padded (T,N) matrix == the reference's per-trajectory rule
    (dagger_trainer.py:199-211) followed by collate_fn's zero padding; collate_fn itself on device tensors."""
    from oracle import vlnce as oracle_vlnce
    from vlnce_b200 import collate

    g = torch.Generator().manual_seed(12)
    batch = []
    for _ in range(37):
        t = int(torch.randint(2, 41, (1,), generator=g))
        a = torch.randint(0, 4, (t,), generator=g)
        prev = torch.cat([torch.zeros(1, dtype=torch.int64), a[:-1]])
        batch.append(({"rgb_features": torch.rand(t, 6, generator=g)}, prev, a, oracle_vlnce.inflection_weights(a, 1.9)))
    want = oracle_vlnce.collate_fn(batch)
    dev_batch = [({k: v.to(DEV) for k, v in o.items()}, p.to(DEV), a.to(DEV), w.to(DEV)) for o, p, a, w in batch]
    got = collate.collate_fn(dev_batch)
    assert torch.equal(got[0]["rgb_features"].cpu(), want[0]["rgb_features"])
    for a, b in zip(got[1:], want[1:]):
        assert a.is_cuda and torch.equal(a.cpu(), b)
    lengths = torch.tensor([b[1].shape[0] for b in batch])
    w = collate.inflection_weights(got[3], 1.9, lengths)
    assert torch.equal(w.cpu(), want[4])
    single = collate.inflection_weights(batch[0][2].to(DEV), 3.2)
    assert torch.equal(single.cpu(), oracle_vlnce.inflection_weights(batch[0][2], 3.2))
