"""trainable conv trunks (``RGB_ENCODER.trainable`` / ``DEPTH_ENCODER.trainable``, config/default.py:242,253).

The taped fp32 forward + layer-by-layer backward of vlnce_b200/trunk.py run on the host models of the C-ABI entry points
(tests/abi_emulator.py) and are compared with
  * the goldens frozen from the unmodified reference (``*_pixels_trunks``: logits, losses, every gradient), and
  * torch autograd through the oracle's trunks evaluated in float64 — the arbiter for how much of the difference to the
    float32 goldens is conditioning: with the seeded synthetic weights the deep trunks' gradients are ill-conditioned
    (train-mode BatchNorm on 4 rows has near-dead channels, rstd ~ 1/sqrt(eps)); torch's own float32 backward sits
    1e-2 away from its float64 one in the early layers, so trunk gradients are held to that yardstick and not to the
    1e-3 of the well-conditioned head."""
import copy
import warnings

import pytest
import torch

import vlnce_b200
from abi_emulator import Emulator
from oracle import cases, vlnce as oracle_vlnce
from test_host_emulated import gold
from vlnce_b200 import lib

warnings.filterwarnings("ignore")


def rel(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-12))


def grouped_errors(got, want):
    groups = {"rgb trunk": [], "depth trunk": [], "head": [], "outputs": []}
    for k, w in want.items():
        if k.startswith("_"):
            continue
        g = got[k].double()
        w = w.double()
        if w.numel() == 11 and k.startswith("grad/"):      # digest of a large tensor: [sum, abs-sum, norm, head[8]]
            e = abs(float(g[2] - w[2])) / max(float(w[2]), 1e-12)
        elif float(w.abs().max()) < 1e-7:
            e = float((g - w).abs().max())
        else:
            e = rel(g, w)
        grp = ("rgb trunk" if "rgb_encoder.cnn" in k else "depth trunk" if "depth_encoder.visual_encoder" in k
               else "head" if k.startswith("grad/") else "outputs")
        groups[grp].append((e, k))
    for v in groups.values():
        v.sort()
    return groups


def check_groups(groups, name):
    for grp, v in groups.items():
        print(f"{name} {grp}: {len(v)} tensors, median {v[len(v) // 2][0]:.2e}, worst {v[-1][0]:.2e} ({v[-1][1]})")
    assert groups["outputs"][-1][0] < 1e-4
    assert groups["head"][-1][0] < 1e-3
    for grp in ("rgb trunk", "depth trunk"):
        assert len(groups[grp]) > 150                                  # every conv / norm parameter has a gradient
        assert groups[grp][len(groups[grp]) // 2][0] < 3e-2 and groups[grp][-1][0] < 6e-2, groups[grp][-3:]


@pytest.mark.parametrize("name", ["cma_train_pixels_trunks", "seq2seq_train_pixels_trunks"])
def test_trainable_trunks_match_reference_golden(name):
    with lib.dry_run(executor=Emulator()) as rec:
        got = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses)
    want = gold(name)
    assert set(got) == {k for k in want if not k.startswith("_")}
    assert rec.calls.count("vlnce_conv2d_ref_wgrad") == 107 and rec.calls.count("vlnce_conv2d_ref_dgrad") == 105
    check_groups(grouped_errors(got, want), name)


def _against_float64(module, run_module, ours: dict):
    def ref(dtype):
        m = copy.deepcopy(module).to(dtype)
        for p in m.parameters():
            p.requires_grad_(True)
            p.grad = None
        run_module(m, dtype)
        return {k: p.grad.double() for k, p in m.named_parameters()}

    r32, r64 = ref(torch.float32), ref(torch.float64)
    worst = (0.0, "")
    for k, t in r64.items():
        mine, torch32 = rel(ours[k], t), rel(r32[k], t)
        assert mine < max(4 * torch32, 2e-3), (k, mine, torch32)
        worst = max(worst, (mine / max(torch32, 1e-9), k))
    return worst


def test_depth_trunk_backward_against_float64_autograd():
    case = cases.CASES["seq2seq_train_pixels_trunks"]
    g0 = torch.Generator().manual_seed(5)
    depth = torch.rand(2, 256, 256, 1, generator=g0)
    up = torch.randn(2, 128, 4, 4, generator=g0)
    enc = cases.build_policy(oracle_vlnce, case).net.depth_encoder.visual_encoder
    denc = cases.build_policy(vlnce_b200, case).net.depth_encoder
    with lib.dry_run(executor=Emulator()):
        out = torch.zeros(2, 16, 128)
        denc.features_into({"depth": depth}, out, 128, "bf16", save=True)   # trainable: runs the fp32 path regardless
        named = dict(denc.visual_encoder.named_parameters())
        g = {k: torch.zeros_like(p) for k, p in named.items()}
        dout = up.flatten(2).permute(0, 2, 1).contiguous()
        denc.backward_trunk(dout.data_ptr(), 128, g, {id(p): k for k, p in named.items()})
    with torch.no_grad():
        assert rel(out, enc({"depth": depth}).flatten(2).permute(0, 2, 1)) < 1e-4   # the taped forward itself
    print("depth trunk: worst ours/torch32 deviation ratio", _against_float64(
        enc, lambda m, dt: m({"depth": depth.to(dt)}).backward(up.to(dt)), g))


@pytest.mark.parametrize("hw,train", [(160, True), (128, False)], ids=["160-train-bn-overlapping-pool", "128-eval-bn"])
def test_rgb_trunk_backward_against_float64_autograd(hw, train):
    case = cases.CASES["cma_train_pixels_trunks"]
    g0 = torch.Generator().manual_seed(6)
    rgb = torch.randint(0, 256, (3, hw, hw, 3), generator=g0, dtype=torch.uint8)
    up = torch.randn(3, 2048, 4, 4, generator=g0)
    cnn = cases.build_policy(oracle_vlnce, case).net.rgb_encoder.cnn
    cnn.train(train)
    renc = cases.build_policy(vlnce_b200, case).net.rgb_encoder
    renc.cnn.train(train)
    with lib.dry_run(executor=Emulator()):
        out = torch.zeros(3, 16, 2048)
        renc.features_into({"rgb": rgb}, out, 2048, "fp32", save=True)
        named = dict(renc.cnn.named_parameters())
        g = {k: torch.zeros_like(p) for k, p in named.items()}
        dout = up.flatten(2).permute(0, 2, 1).contiguous()
        renc.backward_trunk(dout.data_ptr(), 2048, g, {id(p): k for k, p in named.items()})
    print("rgb trunk: worst ours/torch32 deviation ratio", _against_float64(
        cnn, lambda m, dt: m(rgb.to(dt).permute(0, 3, 1, 2) / 255.0).backward(up.to(dt)), g))


def test_pipelined_loop_refuses_trainable_trunks():
    from vlnce_b200.il_step import ILTrainerStep

    case = cases.CASES["cma_train_pixels_trunks"]
    with lib.dry_run():
        pol = cases.build_policy(vlnce_b200, case)
        step = ILTrainerStep(pol, lr=1e-4)
        with pytest.raises(lib.VlnceError):
            step.prefetch_visual({"rgb": torch.zeros(2, 224, 224, 3, dtype=torch.uint8)})


def test_fused_il_step_trains_the_trunks():
    """``ILTrainerStep`` (fused forward + loss + hand-written backward + Adam on the flat parameter buffer) with
    trainable trunks: same losses as the reference golden, the flat gradient covers every trunk tensor and equals the
    autograd-bridge path bit for bit (same kernels), and one Adam step moves the trunk weights."""
    from vlnce_b200.il_step import ILTrainerStep

    name = "seq2seq_train_pixels_trunks"
    case, want = cases.CASES[name], gold(name)
    with lib.dry_run(executor=Emulator()):
        bridge = cases.run_case(vlnce_b200, name, aux=vlnce_b200.AuxLosses)
        pol = cases.build_policy(vlnce_b200, case)
        cases.set_mode(pol, case)
        before = {k: p.detach().clone() for k, p in pol.named_parameters()}
        step = ILTrainerStep(pol, lr=2.5e-4)
        vlnce_b200.AuxLosses.activate()
        loss, action_loss, aux_loss = step.update_agent(*cases.inputs(case))
        vlnce_b200.AuxLosses.deactivate()
    assert abs(loss - float(want["loss"])) < 1e-4 and abs(action_loss - float(want["action_loss"])) < 1e-4
    trunk = [k for k in step.fp.grad_views if "rgb_encoder.cnn" in k or "depth_encoder.visual_encoder" in k]
    assert len(trunk) == 159 + 162
    for k in trunk:
        g = step.fp.grad_views[k].detach()
        b = bridge["grad/" + k]
        assert torch.equal(g if g.numel() <= 4096 else cases.grad_digest(g), b), k
    after = dict(pol.named_parameters())
    k = "net.rgb_encoder.cnn.0.weight"
    assert float((after[k].detach() - before[k]).abs().max()) > 1e-4   # ~lr per element on the first Adam step


def test_trainable_trunk_repacks_weights_after_in_place_update():
    """The fused Adam kernel writes the flat parameter buffer behind torch's version counters; a trainable trunk must
    not keep serving the weights it packed before the update."""
    case = cases.CASES["seq2seq_train_pixels_trunks"]
    g0 = torch.Generator().manual_seed(8)
    depth = torch.rand(2, 256, 256, 1, generator=g0)
    denc = cases.build_policy(vlnce_b200, case).net.depth_encoder
    with lib.dry_run(executor=Emulator()):
        a, b = torch.zeros(2, 16, 128), torch.zeros(2, 16, 128)
        denc.features_into({"depth": depth}, a, 128, "fp32")
        w = denc.visual_encoder.compression[0].weight
        w.data.view(-1).numpy()[::2] *= -1.0        # in place, no version bump (what a device kernel does)
        denc.features_into({"depth": depth}, b, 128, "fp32")
    assert rel(a, b) > 1e-2
