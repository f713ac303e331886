"""Host launch plans executed on host memory through the numpy model of the head kernels
(``tests/abi_emulator.py``) and compared with the golden vectors frozen from the unmodified
reference: pointer arithmetic, leading dimensions, buffer reuse and the hand-written backward are
checked numerically without a GPU.  (The CUDA kernels themselves are checked by the -m gpu tests.)"""
import os
import warnings

import pytest
import torch

import vlnce_b200
from oracle import cases
from vlnce_b200 import lib

from abi_emulator import Emulator

warnings.filterwarnings("ignore")
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gold(name):
    return torch.load(os.path.join(GOLD, name + ".pt"), weights_only=False)


def check(got, want, tol, skip=()):
    worst = ("", 0.0)
    for k, v in got.items():
        if k in skip:
            continue
        g = want[k]
        if v.dtype in (torch.int64, torch.uint8, torch.bool):
            assert torch.equal(v, g), k
            continue
        scale = float(g.abs().max()) + 1e-12
        err = float((v.double() - g.double()).abs().max()) / scale
        if err > worst[1]:
            worst = (k, err)
        assert err < tol or float((v.double() - g.double()).abs().max()) < 1e-7, (k, err)
    return worst


@pytest.mark.parametrize("name", ["cma_train_features", "cma_train_features_nopm", "seq2seq_train_features",
                                  "cma_rxr_train_features", "cma_train_features_abl_i", "cma_train_features_abl_v",
                                  "seq2seq_train_features_abl", "cma_train_features_lstm",
                                  "seq2seq_train_features_lstm", "cma_train_features_igru",
                                  "seq2seq_train_features_igru"])
def test_il_train_plan_matches_golden(name):
    with lib.dry_run(executor=Emulator()):
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses)
    worst = check(got, gold(name), 2e-4)
    print(name, "worst", worst)


def test_rxr_instruction_projection_takes_the_tensor_core_route(monkeypatch):
    """Reduced-precision modes put the RxR instruction-feature input projection and its weight gradient on the TF32
    tensor-core tiles once the batch has enough token rows (engine.HEAD_TC_MIN_ROWS): rounded operands, one
    ``vlnce_conv_tma`` (1x1 mode = GEMM, plain fp32 result) per direction forward, transposed token-minor operands and
    ONE stacked launch for both directions' weight gradients.  The host model multiplies in exact fp32, so this checks
    the routing, the transposes and the gradient placement; precision="fp32" must stay on the CUDA-core GEMM."""
    from vlnce_b200 import engine

    monkeypatch.setattr(engine, "HEAD_TC_MIN_ROWS", 64)
    name = "cma_rxr_train_features"        # 9 rows x 40 tokens = 360 token rows... not a multiple of 32: forward only
    with lib.dry_run(executor=Emulator()) as rec:
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="bf16")
    assert rec.calls.count("vlnce_conv_tma") == 2 and rec.calls.count("vlnce_round_tf32") == 3
    assert "vlnce_transpose_f32" not in rec.calls
    print(name, "tensor-core projection worst", check(got, gold(name), 2e-4))
    with lib.dry_run(executor=Emulator()) as rec:
        cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="fp32")
    assert "vlnce_conv_tma" not in rec.calls
    # a batch whose token-row count is a multiple of 32 (8 rows x 40 tokens) also takes the weight-gradient route:
    # same gradients as the CUDA-core plan of the same batch
    monkeypatch.setitem(cases.CASES, "_rxr_tc", dict(cases.CASES[name], t=2, n=4))
    with lib.dry_run(executor=Emulator()) as rec:
        tc = cases.run_case(vlnce_b200, "_rxr_tc", aux=vlnce_b200.AuxLosses, precision="bf16")
    assert rec.calls.count("vlnce_conv_tma") == 3 and rec.calls.count("vlnce_transpose_f32") == 3
    monkeypatch.setattr(engine, "HEAD_TC_MIN_ROWS", 0)
    with lib.dry_run(executor=Emulator()):
        simt = cases.run_case(vlnce_b200, "_rxr_tc", aux=vlnce_b200.AuxLosses, precision="bf16")
    print("tensor-core weight-gradient route vs CUDA-core plan worst", check(tc, simt, 2e-5))


# ---------------------------------------------------------------------------------------------
# waypoint path: the conv trunks are not emulated — their features come from the oracle's trunks
# ---------------------------------------------------------------------------------------------
def substitute_trunks(pol, case):
    from oracle import waypoint as oracle_wp

    opol = cases.build_policy(oracle_wp, case)
    cases.set_mode(opol, case)

    def rgb_into(observations, out, ld, precision):
        x = observations["rgb"].float().permute(0, 3, 1, 2) / 255.0
        with torch.no_grad():
            f = opol.net.rgb_encoder.cnn(x)[:, :, 0, 0]         # (F, 512): the global pool, broadcast 4x4 (Q2)
        out.view(out.shape[0], -1)[:, : f.shape[1]] = f          # position 0, channels [0, 512)

    def dep_into(observations, out, ld, precision):
        with torch.no_grad():
            f = opol.net.depth_encoder.visual_encoder({"depth": observations["depth"]})  # (F,128,4,4)
        out[:, :, : f.shape[1]] = f.flatten(2).permute(0, 2, 1)

    pol.net.rgb_encoder.features_into = rgb_into
    pol.net.depth_encoder.features_into = dep_into


class _Impl:
    """vlnce_b200 with the trunks of every built WaypointPolicy substituted (test only)."""

    AuxLosses = vlnce_b200.AuxLosses

    def __init__(self, case):
        self.case = case

    @property
    def WaypointPolicy(self):
        case = self.case

        class Wrapped:
            @staticmethod
            def from_config(*a, **kw):
                pol = vlnce_b200.WaypointPolicy.from_config(*a, **kw)
                substitute_trunks(pol, case)
                return pol

        return Wrapped


@pytest.mark.parametrize("name", ["wp_act", "wp_act_u8", "wp_eval", "wp_act_hpn", "wp_eval_hpn", "wp_act_dc", "wp_act_dd",
                                  "wp_eval_dd", "wp_eval_d_", "wp_eval_abl", "wp_eval_lstm", "wp_eval_igru"])
def test_waypoint_plan_matches_golden(name):
    case = cases.CASES[name]
    with lib.dry_run(executor=Emulator()):
        got = cases.run_case(_Impl(case), name, aux=vlnce_b200.AuxLosses)
    want = gold(name)
    assert set(got) == {k for k in want if not k.startswith("_")}
    worst = check(got, want, 3e-4)
    print(name, "worst", worst)


def _fill_rollouts(st, T, N, length=12, classes=None, layers=2):
    from vlnce_b200 import synth

    g = torch.Generator().manual_seed(9)
    first = synth.make_waypoint_rows(g, N, length=length)
    for k in st.observations:
        st.observations[k][0].copy_(first[k])
    for t in range(T):
        a = synth.make_waypoint_actions(g, N, classes=classes)
        st.insert(synth.make_waypoint_rows(g, N, length=length), torch.randn(N, layers, 256, generator=g) * 0.3, a,
                  -2 + 0.2 * torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g) * 0.3,
                  torch.randn(N, 1, generator=g), torch.ones(N, 1, dtype=torch.uint8))
    return torch.randn(N, 1, generator=g) * 0.3


@pytest.mark.parametrize("name", ["wp_eval", "wp_eval_dd", "wp_eval_igru"])
def test_fused_wddppo_update_matches_oracle_loop(name):
    """``WDDPPOStep.update`` (fused loss+gradient kernel, hand-written backward, clip, Adam; device-style rollout
    storage with the GAE kernel) against the restated ``WDDPPO.update`` loop (itself pinned to the reference's
    class in tests/test_waypoint_oracle.py): same six scalars, same updated parameters."""
    from oracle import waypoint as oracle_wp
    from vlnce_b200.ppo_step import ActionDictRolloutStorage, WDDPPOStep

    case = cases.CASES[name]
    T, N = 2, 2
    Lr = cases.layers_of(case)  # 2 GRU slots, or 4 with LSTM state encoders (wp_eval_igru: GRU instruction encoder too)
    kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, offset_regularize_coef=0.1146,
              pano_entropy_coef=1.0, offset_entropy_coef=1.0, distance_entropy_coef=1.0)
    space = oracle_wp.storage_obs_space(length=12)
    wp = case.get("waypoint", {})
    cont = dict(continuous_offset=wp.get("continuous_offset", True), continuous_distance=wp.get("continuous_distance", True))
    classes = dict(distance=0 if cont["continuous_distance"] else 6, offset=0 if cont["continuous_offset"] else 7)
    # oracle
    opol = cases.build_policy(oracle_wp, case)
    cases.set_mode(opol, case)
    ost = oracle_wp.ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=Lr, **cont)
    nv = _fill_rollouts(ost, T, N, classes=classes, layers=Lr)
    ost.compute_returns(nv, True, 0.99, 0.95)
    opt = torch.optim.Adam([p for p in opol.parameters() if p.requires_grad], lr=2e-4, eps=1e-5)
    torch.manual_seed(21)
    want = oracle_wp.ppo_update(opol, opt, ost, ppo_epoch=1, num_mini_batch=2, max_grad_norm=0.2,
                                use_clipped_value_loss=True, **kw)
    # fused path on the emulator
    pol = cases.build_policy(vlnce_b200, case)
    cases.set_mode(pol, case)
    substitute_trunks(pol, case)
    st = ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=Lr, **cont)
    nv2 = _fill_rollouts(st, T, N, classes=classes, layers=Lr)
    with lib.dry_run(executor=Emulator()):
        st.compute_returns(nv2, True, 0.99, 0.95)
        torch.testing.assert_close(st.returns, ost.returns, rtol=1e-6, atol=1e-6)
        agent = WDDPPOStep(pol, ppo_epoch=1, num_mini_batch=2, lr=2e-4, eps=1e-5, max_grad_norm=0.2,
                           use_clipped_value_loss=True, **kw)
        torch.manual_seed(21)
        got = agent.update(st)
    torch.testing.assert_close(torch.tensor(got), torch.tensor(want), rtol=2e-4, atol=2e-6)
    ref = dict(opol.named_parameters())
    for k, p in pol.named_parameters():
        if p.requires_grad:
            # Adam's first steps move every weight by ~lr regardless of the gradient's size: compare the UPDATE
            torch.testing.assert_close(p.detach(), ref[k].detach(), rtol=0, atol=4e-5, msg=lambda m: f"{k}: {m}")
    if name != "wp_eval":
        return
    # the same update fed the frozen trunks' cached features instead of the stored frames (enable_visual_cache): what a
    # collection loop stores between act() and insert() is reproduced here with get_value on every stored step
    from vlnce_b200.waypoint_policy import VISUAL_CACHE_KEYS

    pol2 = cases.build_policy(vlnce_b200, case)
    cases.set_mode(pol2, case)
    substitute_trunks(pol2, case)
    st2 = ActionDictRolloutStorage(T, N, space, 256, num_recurrent_layers=Lr, **cont)
    nv3 = _fill_rollouts(st2, T, N, classes=classes, layers=Lr)
    with lib.dry_run(executor=Emulator()) as rec:
        st2.compute_returns(nv3, True, 0.99, 0.95)
        st2.enable_visual_cache(pol2)
        for s_ in range(T):
            obs = {k: v[s_] for k, v in st2.observations.items() if k not in VISUAL_CACHE_KEYS}
            pol2.get_value(obs, st2.recurrent_hidden_states[s_], {k: v[s_] for k, v in st2.prev_actions.items()},
                           st2.masks[s_])
            st2.step = s_
            st2.store_visual_features(pol2.visual_features())
        st2.step = T
        pol2.net.rgb_encoder.features_into = pol2.net.depth_encoder.features_into = None   # the update must not need them
        n_before = len(rec.calls)
        agent2 = WDDPPOStep(pol2, ppo_epoch=1, num_mini_batch=2, lr=2e-4, eps=1e-5, max_grad_norm=0.2,
                            use_clipped_value_loss=True, **kw)
        torch.manual_seed(21)
        got2 = agent2.update(st2)
        assert "vlnce_wp_stack_frames" not in rec.calls[n_before:]
    torch.testing.assert_close(torch.tensor(got2), torch.tensor(want), rtol=2e-4, atol=2e-6)
    for k, p in pol2.named_parameters():
        if p.requires_grad:
            torch.testing.assert_close(p.detach(), ref[k].detach(), rtol=0, atol=4e-5, msg=lambda m: f"cached {k}: {m}")


# ---------------------------------------------------------------------------------------------
# fused imitation-learning step (the benchmarked path): ILTrainerStep.update_agent on the emulator against the
# restated ``_update_agent`` + torch Adam (common/base_il_trainer.py:134-180)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,accum", [("cma_train_features", 1), ("cma_rxr_train_features", 1),
                                        ("seq2seq_train_features", 1), ("cma_train_features_nopm", 2),
                                        ("cma_train_features_lstm", 1), ("seq2seq_train_features_igru", 1)])
def test_fused_il_step_matches_oracle_update(name, accum):
    from oracle import vlnce as oracle_vlnce
    from vlnce_b200.il_step import ILTrainerStep

    case = cases.CASES[name]
    impl = cases.oracle_impl(case)
    opol = cases.build_policy(impl, case)
    cases.set_mode(opol, case)
    opt = torch.optim.Adam([p for p in opol.parameters() if p.requires_grad], lr=2.5e-4)
    aux = oracle_vlnce.AuxLosses
    aux.activate()
    batch = cases.inputs(case)
    want = []
    for k in range(accum):
        want.append(oracle_vlnce.update_agent(opol, opt, *batch, step_grad=(k == accum - 1),
                                              loss_accumulation_scalar=accum, aux=aux))
    aux.deactivate()

    pol = cases.build_policy(vlnce_b200, case)
    cases.set_mode(pol, case)
    vlnce_b200.AuxLosses.activate()
    step = ILTrainerStep(pol, lr=2.5e-4)
    got = []
    with lib.dry_run(executor=Emulator()):
        for k in range(accum):
            got.append(step.update_agent(*batch, step_grad=(k == accum - 1), loss_accumulation_scalar=accum))
    vlnce_b200.AuxLosses.deactivate()
    torch.testing.assert_close(torch.tensor(got), torch.tensor(want), rtol=2e-5, atol=1e-6)
    ref = dict(opol.named_parameters())
    moved = 0
    for k, p in pol.named_parameters():
        if p.requires_grad:
            # first Adam step: |update| ~ lr wherever the gradient is not tiny; a wrong sign shows as 2*lr
            torch.testing.assert_close(p.detach(), ref[k].detach(), rtol=0, atol=5e-5, msg=lambda m: f"{k}: {m}")
            moved += 1
    assert moved >= 10


def test_pipelined_loop_plan_equals_plain_loop():
    """Host logic of the pipelined training loop (ILTrainerStep.pipelined_step: visual features of batch i+1 into
    the 'next' slot while batch i trains from the 'cur' slot, then rotate) on the emulator with cached features:
    bit-identical losses and parameters to the plain sequential loop (the emulator is deterministic)."""
    from vlnce_b200 import synth
    from vlnce_b200.il_step import ILTrainerStep

    case = dict(cases.CASES["cma_train_features"], t=2, n=3)
    batches = [synth.make_train_batch(700 + i, 2, 3, length=16, features=True) for i in range(3)]
    VIS = ("rgb_features", "depth_features")

    def make():
        pol = cases.build_policy(vlnce_b200, case).eval()
        return pol, ILTrainerStep(pol, lr=1e-3)

    aux = vlnce_b200.AuxLosses
    aux.activate()
    with lib.dry_run(executor=Emulator()):
        _, st_a = make()
        plain = [st_a.update_agent(*b) for b in batches]
        _, st_b = make()
        st_b._prepare()
        st_b.prefetch_visual(batches[0][0])
        st_b.rotate_visual(batches[0][1].shape[0])
        piped = []
        for i, b in enumerate(batches):
            nxt = {k: batches[(i + 1) % len(batches)][0][k] for k in VIS}
            obs_nv = {k: v for k, v in b[0].items() if k not in VIS}
            piped.append(st_b.pipelined_step(obs_nv, b[1], b[2], b[3], b[4], nxt).tolist())
    aux.deactivate()
    torch.testing.assert_close(torch.tensor(piped), torch.tensor(plain), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(st_b.fp.flat, st_a.fp.flat, rtol=0, atol=1e-7)


@pytest.mark.parametrize("base,t,n,length", [("cma_train_features", 1, 1, 3), ("cma_train_features", 7, 1, 80),
                                             ("seq2seq_train_features", 1, 2, 1), ("cma_rxr_train_features", 2, 1, 5)])
def test_edge_shapes_match_oracle(base, t, n, length):
    """Smallest / most lopsided batches (one row, one trajectory, one-token instructions): the launch plan against
    the oracle run live on the same seeded inputs."""
    name = f"_edge_{base}_{t}_{n}_{length}"
    cases.CASES[name] = dict(cases.CASES[base], t=t, n=n, length=length)
    try:
        want = cases.run_case(cases.oracle_impl(cases.CASES[name]), name, aux=__import__("oracle.vlnce", fromlist=["x"]).AuxLosses)
        with lib.dry_run(executor=Emulator()):
            got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses)
    finally:
        del cases.CASES[name]
    assert set(got) == set(want)
    print(name, "worst", check(got, want, 3e-4))


@pytest.mark.parametrize("base,over", [("wp_eval", dict(t=1, n=1, length=3)), ("wp_act", dict(n=1, length=2)),
                                       ("wp_eval_dd", dict(t=4, n=1, length=9))])
def test_waypoint_edge_shapes_match_oracle(base, over):
    from oracle import waypoint as oracle_wp

    name = "_edge_" + base + "_".join(f"{k}{v}" for k, v in over.items())
    cases.CASES[name] = dict(cases.CASES[base], **over)
    try:
        case = cases.CASES[name]
        want = cases.run_case(oracle_wp, name)
        with lib.dry_run(executor=Emulator()):
            got = cases.run_case(_Impl(case), name, aux=vlnce_b200.AuxLosses)
    finally:
        del cases.CASES[name]
    assert set(got) == set(want)
    print(name, "worst", check(got, want, 3e-4))


def test_recollect_style_loop_matches_oracle():
    """BASELINE config C3 shape family at reduced size: the RecollectTrainer teacher-forcing loop
    (``recollect_trainer.py:128-185``) — ragged trajectories -> collate_fn -> ``_update_agent`` with gradient
    accumulation (``loss_accumulation_scalar = effective_batch_size // batch_size`` and the reference's
    ``step_grad = bool(step_id % scalar)`` pattern).  Product side: device-style collate + inflection-weight kernel +
    fused ILTrainerStep on the emulator; oracle side: the restated collate_fn / update_agent with torch Adam."""
    from oracle import vlnce as oracle_vlnce
    from vlnce_b200 import collate
    from vlnce_b200.il_step import ILTrainerStep

    case = dict(cases.CASES["cma_train_features"], pm=True)
    g = torch.Generator().manual_seed(77)

    def trajectory():
        t = int(torch.randint(2, 7, (1,), generator=g))
        obs = {"rgb_features": torch.rand(t, 2048, 4, 4, generator=g), "depth_features": torch.rand(t, 128, 4, 4, generator=g),
               "instruction": torch.randint(1, 2504, (1, 12), generator=g).expand(t, 12).contiguous(),
               "progress": torch.rand(t, 1, generator=g)}
        actions = torch.randint(0, 4, (t,), generator=g)
        prev = torch.cat([torch.zeros(1, dtype=torch.int64), actions[:-1]])
        return obs, prev, actions
    batches = [[trajectory() for _ in range(3)] for _ in range(4)]   # 4 micro-batches of 3 trajectories
    scalar = 2                                                       # effective_batch_size 6 // batch_size 3

    opol = cases.build_policy(oracle_vlnce, case).eval()
    opt = torch.optim.Adam([p for p in opol.parameters() if p.requires_grad], lr=2.5e-4)
    oracle_vlnce.AuxLosses.activate()
    want = []
    for step_id, trajs in enumerate(batches):
        b = oracle_vlnce.collate_fn([(o, p, a, oracle_vlnce.inflection_weights(a, 3.2)) for o, p, a in trajs])
        want.append(oracle_vlnce.update_agent(opol, opt, *b, step_grad=bool(step_id % scalar),
                                              loss_accumulation_scalar=scalar, aux=oracle_vlnce.AuxLosses))
    oracle_vlnce.AuxLosses.deactivate()

    pol = cases.build_policy(vlnce_b200, case).eval()
    vlnce_b200.AuxLosses.activate()
    step = ILTrainerStep(pol, lr=2.5e-4)
    got = []
    with lib.dry_run(executor=Emulator()):
        for step_id, trajs in enumerate(batches):
            lengths = torch.tensor([a.shape[0] for _, _, a in trajs])
            obs, prev, masks, actions, _ = collate.collate_fn([(o, p, a, torch.zeros(a.shape[0])) for o, p, a in trajs])
            weights = collate.inflection_weights(actions, 3.2, lengths)   # one launch on the padded (T,N) matrix
            got.append(step.update_agent(obs, prev, masks, actions, weights, step_grad=bool(step_id % scalar),
                                         loss_accumulation_scalar=scalar))
    vlnce_b200.AuxLosses.deactivate()
    torch.testing.assert_close(torch.tensor(got), torch.tensor(want), rtol=3e-5, atol=1e-6)
    ref = dict(opol.named_parameters())
    for k, p in pol.named_parameters():
        if p.requires_grad:
            torch.testing.assert_close(p.detach(), ref[k].detach(), rtol=0, atol=1e-4, msg=lambda m: f"{k}: {m}")


def test_update_agent_from_host_equals_update_agent():
    """``ILTrainerStep.update_agent_from_host`` stages the batch into persistent buffers on the policy's device and runs
    the same fused step (here both live on the host model's memory)."""
    from vlnce_b200.il_step import ILTrainerStep

    case = cases.CASES["cma_train_features"]
    batch = cases.inputs(case)
    res = []
    for from_host in (False, True):
        pol = cases.build_policy(vlnce_b200, case).eval()
        vlnce_b200.AuxLosses.activate()
        step = ILTrainerStep(pol, lr=2.5e-4)
        with lib.dry_run(executor=Emulator()):
            fn = step.update_agent_from_host if from_host else step.update_agent
            out = [fn(*batch) for _ in range(2)]           # second call re-uses the staging buffers
        vlnce_b200.AuxLosses.deactivate()
        res.append((out, step.fp.flat.clone()))
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1])
    assert len(step._host_stage) == len(batch[0]) + 4


def test_padded_rows_bypass_the_trunks_without_changing_the_update():
    """``ILTrainerStep.skip_padded_rows``: rows of a ragged teacher-forcing batch whose loss weight is 0 (padding after
    the end of a trajectory) skip the frozen eval-mode conv trunks — same losses, same gradients as computing them."""
    from vlnce_b200.il_step import ILTrainerStep

    case = cases.CASES["cma_train_pixels"]                  # t = 3, n = 2, raw 256x256 frames
    obs, prev, masks, actions, weights = cases.inputs(case)
    weights = weights.clone()
    weights[2, 1] = 0.0                                      # trajectory 1 ended after two steps
    weights[1:, 0] = 0.0                                     # trajectory 0 after one
    res = []
    for skip in (False, True):
        pol = cases.build_policy(vlnce_b200, case).eval()
        pol.net.precision = "fp32"
        vlnce_b200.AuxLosses.activate()
        step = ILTrainerStep(pol, lr=2.5e-4)
        step.skip_padded_rows = skip
        with lib.dry_run(executor=Emulator()) as rec:
            step.compute_grads(obs, prev, masks, actions, weights)
        vlnce_b200.AuxLosses.deactivate()
        res.append((step._losses.clone(), step.fp.grad.clone(), rec.calls.count("vlnce_conv2d_ref")))
    (l0, g0, n0), (l1, g1, n1) = res
    assert n0 == n1                                          # same layers ... on half the images
    torch.testing.assert_close(l1, l0, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(g1, g0, rtol=1e-5, atol=1e-8)
    assert float(g0.abs().max()) > 0


def test_ddppo_collect_and_update_loop_matches_oracle():
    """BASELINE config C4 shape family at reduced size: the DD-PPO waypoint loop of
    ``ddppo_waypoint_trainer.py:140-300,470-560`` with a synthetic environment — ``act`` (deterministic, so both
    sides take the same actions) -> history-frame gather -> ``rollouts.insert`` -> ``get_value`` -> GAE ->
    ``WDDPPO.update``.  Product: WaypointPolicy / device storage / fused WDDPPOStep on the emulator; oracle: the
    restated policy, storage and update loop."""
    from oracle import waypoint as oracle_wp
    from vlnce_b200 import synth
    from vlnce_b200.ppo_step import ActionDictRolloutStorage, WDDPPOStep, select_history_frames

    case = cases.CASES["wp_eval"]
    T, N, L = 2, 2, 10
    space = oracle_wp.storage_obs_space(length=L)
    g = torch.Generator().manual_seed(31)
    first = synth.make_waypoint_rows(g, N, length=L)
    nxt = [synth.make_waypoint_rows(g, N, length=L) for _ in range(T)]
    rewards = [torch.randn(N, 1, generator=g) for _ in range(T)]
    dones = [torch.tensor([[0], [1]], dtype=torch.uint8), torch.tensor([[1], [1]], dtype=torch.uint8)]  # masks = not done
    kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, offset_regularize_coef=0.1146,
              pano_entropy_coef=1.0, offset_entropy_coef=1.0, distance_entropy_coef=1.0)

    def run(pol, storage_cls, update):
        st = storage_cls(T, N, space, 256, num_recurrent_layers=2)
        for k in st.observations:
            st.observations[k][0].copy_(first[k])
        for t in range(T):
            obs = {k: v[st.step] for k, v in st.observations.items()}
            prev = {k: v[st.step].clone() for k, v in st.prev_actions.items()}
            with torch.no_grad():
                value, actions, elems, _, _, lp, h, _ = pol.act(obs, st.recurrent_hidden_states[st.step], prev,
                                                                st.masks[st.step], deterministic=True)
            hist = select_history_frames(obs, elems["pano"], 12)
            batch = dict(nxt[t], rgb_history=hist["rgb"], depth_history=hist["depth"])
            st.insert(batch, h, elems, lp, value, rewards[t], dones[t])
        with torch.no_grad():
            nv = pol.get_value({k: v[st.step] for k, v in st.observations.items()}, st.recurrent_hidden_states[st.step],
                               {k: v[st.step].clone() for k, v in st.prev_actions.items()}, st.masks[st.step])
        st.compute_returns(nv, True, 0.99, 0.95)
        torch.manual_seed(5)
        return st, update(pol, st)

    opol = cases.build_policy(oracle_wp, case)
    cases.set_mode(opol, case)
    opt = torch.optim.Adam([p for p in opol.parameters() if p.requires_grad], lr=2e-4, eps=1e-5)
    ost, want = run(opol, oracle_wp.ActionDictRolloutStorage,
                    lambda p, s: oracle_wp.ppo_update(p, opt, s, ppo_epoch=1, num_mini_batch=2, max_grad_norm=0.2,
                                                      use_clipped_value_loss=True, **kw))
    pol = cases.build_policy(vlnce_b200, case)
    cases.set_mode(pol, case)
    substitute_trunks(pol, case)
    with lib.dry_run(executor=Emulator()):
        agent = WDDPPOStep(pol, ppo_epoch=1, num_mini_batch=2, lr=2e-4, eps=1e-5, max_grad_norm=0.2,
                           use_clipped_value_loss=True, **kw)
        st, got = run(pol, ActionDictRolloutStorage, lambda p, s: agent.update(s))
    for name in ("returns", "value_preds", "action_log_probs", "rewards", "recurrent_hidden_states"):
        torch.testing.assert_close(getattr(st, name), getattr(ost, name), rtol=1e-4, atol=1e-5, msg=lambda m: f"{name}: {m}")
    for k in st.actions:
        torch.testing.assert_close(st.actions[k], ost.actions[k], rtol=1e-4, atol=1e-5)
    assert torch.equal(st.observations["rgb_history"], ost.observations["rgb_history"])
    torch.testing.assert_close(torch.tensor(got), torch.tensor(want), rtol=3e-4, atol=3e-6)
    ref = dict(opol.named_parameters())
    for k, p in pol.named_parameters():
        if p.requires_grad:
            torch.testing.assert_close(p.detach(), ref[k].detach(), rtol=0, atol=5e-5, msg=lambda m: f"{k}: {m}")


# ---------------------------------------------------------------------------------------------
# raw-pixel paths in the fp32 precision mode: the conv-trunk launch plans (trunk.py) on the emulator as well
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cma_act", "cma_act_bntrain", "cma_act_rgb224", "seq2seq_act", "cma_train_pixels",
                                  "cma_act_lstm", "cma_act_nrgb"])
def test_pixel_paths_fp32_match_golden(name):
    """Both conv trunks from raw pixels in ``precision="fp32"`` (vlnce_conv2d_ref + GroupNorm / BatchNorm / pooling
    kernels, modelled with torch CPU ops): block wiring, residual and downsample routing, folded eval-mode and
    train-mode BatchNorm, the 4x4 / global pools (Q2) against the reference goldens."""
    with lib.dry_run(executor=Emulator()):
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="fp32")
    print(name, "worst", check(got, gold(name), 3e-4))


@pytest.mark.parametrize("name", ["cma_act", "seq2seq_act", "cma_act_rgb224"])
def test_pixel_paths_tf32_routing_matches_golden(name, monkeypatch):
    """``precision="tf32"``: fp32 storage, every stride-1 1x1 / 3x3 and stride-2 layer the persistent TMA kernel covers
    goes through ``vlnce_conv_tma`` / ``vlnce_conv_s2_tma`` with fp32 operands, the RGB stem through the fp16 s2d path
    with fp32 output.  The host model multiplies in exact fp32 but sees the same packed operands as the GPU (weights
    rounded to TF32 once, fp16 stem weights), so this checks the routing, the fp32 residual path and the per-layer
    GroupNorm depth path at the mode's own tolerance (measured on the B200: hidden state 7e-4..1e-3)."""
    from vlnce_b200.encoders import resnet_encoders

    # (the default keeps the depth trunk on fp16 operands in this mode — same mantissa width; here: the per-layer
    # fp32-storage path for both trunks)
    monkeypatch.setattr(resnet_encoders, "TF32_DEPTH_FP16", False)
    with lib.dry_run(executor=Emulator()) as rec:
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="tf32")
    assert rec.calls.count("vlnce_conv_tma") > 60 and "vlnce_conv_stem_tma" in rec.calls
    # what the TMA tiles do not cover at this batch (4x4 images, image counts that do not fill a tile) runs on the
    # warp-cooperative fp32 conv, never on the one-thread-per-output reference conv
    assert "vlnce_conv2d_warp" in rec.calls and "vlnce_conv2d_ref" not in rec.calls
    if name != "cma_act_rgb224":
        assert "vlnce_conv_s2_tma" in rec.calls
    print(name, "tf32 routing worst", check(got, gold(name), 3e-3))
    monkeypatch.setattr(resnet_encoders, "TF32_DEPTH_FP16", True)
    with lib.dry_run(executor=Emulator()) as rec:
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="tf32")
    assert "vlnce_conv2d_bf16" in rec.calls or "vlnce_depth_fused_run" in rec.calls   # depth trunk: 16-bit (fp16) tiles
    print(name, "tf32 routing with the fp16 depth trunk worst", check(got, gold(name), 1e-2))


@pytest.mark.parametrize("name", ["wp_act", "wp_act_nrgb"])
def test_waypoint_pixel_path_fp32_matches_golden(name):
    """WaypointPolicy.act from raw pano pixels (ResNet-18 + DD-PPO ResNet-50 over the stacked frames), no trunk
    substitution; ``wp_act_nrgb``: uint8 frames with MODEL.normalize_rgb."""
    with lib.dry_run(executor=Emulator()) as rec:
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="fp32")
    assert ("vlnce_rgb_center" in rec.calls) == name.endswith("nrgb")
    print(name, "pixels worst", check(got, gold(name), 3e-4))


@pytest.mark.parametrize("name,tol", [("cma_act", 1e-2), ("cma_act_rgb224", 1e-2), ("cma_act_bntrain", 2e-2),
                                      ("seq2seq_act", 1e-2), ("wp_act", 2e-2), ("cma_act_nrgb", 1e-2),
                                      ("wp_act_nrgb", 2e-2)])
def test_pixel_paths_16bit_packing_matches_golden(name, tol, monkeypatch):
    """The default 16-bit trunk paths on functional models of the tensor-core entry points: what is checked on the
    host is trunk.py's operand packing and routing — KRSC weights, the 2x2 space-to-depth RGB stem (cell image +
    re-indexed 4x4 weights), the padded im2col depth stem (gemm_c), fp16 vs bf16 trunk dtypes, folded / batch-stat
    norms — at the 16-bit tolerance of BASELINE.json (outputs rounded to bf16 / fp16 after every layer as on the GPU)."""
    from vlnce_b200 import trunk as trunk_mod

    # the committed B200 tile table picks a measured kernel per geometry; this host check wants the un-tuned default
    # routing (persistent TMA tiles wherever they are legal), so the table is emptied for the walk
    monkeypatch.setattr(trunk_mod, "_TUNE_CACHE", {})
    want = gold(name)
    with lib.dry_run(executor=Emulator()) as rec:
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="bf16")
    assert "vlnce_stem_s2d" in rec.calls and "vlnce_conv_stem_tma" in rec.calls and "vlnce_im2col_depth_stem" in rec.calls
    assert rec.calls.count("vlnce_conv_tma") > 20
    logits = "pano_stop_logits" if name.startswith("wp_") else "logits"
    e_p = float((got[logits].exp() - want[logits].exp()).abs().max() / want[logits].exp().abs().max())
    e_h = float((got["rnn_states_out"] - want["rnn_states_out"]).norm() / want["rnn_states_out"].norm())
    print(f"{name}[16-bit, host model] rel err probs {e_p:.3e} hidden {e_h:.3e}")
    assert e_p < tol and e_h < tol


# ---------------------------------------------------------------------------------------------
# size-independent properties of the path (run at BASELINE's full size on the GPU, tests/test_zz_gpu_variants.py)
# ---------------------------------------------------------------------------------------------
def path_properties(pol, n, hw, length, device, tol, rgb_hw=None):
    """Properties the reference's arithmetic guarantees for ``act`` in eval mode, checked on ``pol``:
    rows are independent (a batch equals its halves), repeated calls agree, zero-padding the instruction further does
    not change anything (packed LSTM + masked attention), and ``masks == 0`` rows ignore the carried hidden state and
    the previous action (``rnn_state_encoder`` reset, ``((a + 1) * mask)`` index).  Returns the worst deviation."""
    from vlnce_b200 import synth

    obs, h, prev, masks = synth.make_act_batch(77, n, hw=hw, rgb_hw=rgb_hw, length=length)
    masks[: n // 4] = 0
    to = lambda t: t.to(device)  # noqa: E731
    obs, h, prev, masks = {k: to(v) for k, v in obs.items()}, to(h), to(prev), to(masks)

    def run(o, hh, pp, mm):
        with torch.no_grad():
            d = pol.build_distribution(o, hh, pp, mm)
            _, h_out = pol.act(o, hh, pp, mm, deterministic=True)
        return d.logits.exp().float().cpu(), h_out.float().cpu()

    def dev(a, b):
        return max(float((a[0] - b[0]).abs().max()), float((a[1] - b[1]).norm() / b[1].norm()))

    full = run(obs, h, prev, masks)
    worst = {"repeat": dev(run(obs, h, prev, masks), full)}
    half = n // 2
    parts = [run({k: v[s] for k, v in obs.items()}, h[s], prev[s], masks[s]) for s in (slice(0, half), slice(half, n))]
    worst["halves"] = dev((torch.cat([p[0] for p in parts]), torch.cat([p[1] for p in parts])), full)
    longer = dict(obs, instruction=torch.cat([obs["instruction"], torch.zeros_like(obs["instruction"][:, :9])], 1))
    worst["padding"] = dev(run(longer, h, prev, masks), full)
    h2, prev2 = h.clone(), prev.clone()
    h2[: n // 4] = 3.0
    prev2[: n // 4] = (prev2[: n // 4] + 1) % 4
    worst["mask reset"] = dev(run(obs, h2, prev2, masks), full)
    assert torch.isfinite(full[0]).all() and abs(float(full[0].sum(1).mean()) - 1.0) < 1e-3
    for k, v in worst.items():
        assert v <= tol, (k, v, worst)
    return worst


def test_path_properties_small():
    case = dict(cases.CASES["cma_act"], n=4)
    with lib.dry_run(executor=Emulator()):
        pol = cases.build_policy(vlnce_b200, case)
        pol.net.precision = "fp32"
        pol.eval()
        print("properties (host model, 4 rows):", path_properties(pol, 4, 256, 24, "cpu", 1e-5, rgb_hw=64))


@pytest.mark.parametrize("name", ["cma_train_pixels", "seq2seq_train_pixels"])
def test_pixel_training_16bit_packing_matches_golden(name):
    """The training step from raw pixels on the 16-bit trunk models: outputs / losses at the path's 16-bit tolerance,
    head gradients by relative L2 (same yardstick as tests/test_gpu_train.py's bf16 cases)."""
    from test_trunk_backward import grouped_errors

    with lib.dry_run(executor=Emulator()):
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses, precision="bf16")
    g = grouped_errors(got, gold(name))
    assert g["outputs"][-1][0] < 1e-2
    head = g["head"]
    print(name, "[16-bit, host model] outputs %.2e, head gradients median %.2e worst %.2e (%s)"
          % (g["outputs"][-1][0], head[len(head) // 2][0], head[-1][0], head[-1][1]))
    assert head[len(head) // 2][0] < 5e-3 and head[-1][0] < 0.1
