"""Waypoint path: the restatement in oracle/waypoint.py against the UNMODIFIED
reference (build container) and against frozen fixtures (everywhere)."""
import math
import os
import warnings

import pytest
import torch

from oracle import reference_import, waypoint

warnings.filterwarnings("ignore")
GOLD = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(not reference_import.available(), reason="/root/reference not present")


def tn_inputs():
    g = torch.Generator().manual_seed(7)
    loc = 0.25 + torch.rand(64, 1, generator=g) * 3.75
    scale = torch.sqrt(0.01 + torch.rand(64, 1, generator=g) * 3.5)
    value = 0.25 + torch.rand(64, 1, generator=g) * 3.75
    return loc, scale, value


def tn_outputs(cls):
    loc, scale, value = tn_inputs()
    out = {}
    for tag, (lo, hi, l, s, v) in {
            "distance": (0.25, 4.0, loc, scale, value),
            "offset": (-math.pi / 12, math.pi / 12, (loc - 2.1) / 8.0, scale * 0.1, (value - 2.1) / 8.0)}.items():
        d = cls(l, s, lo, hi)
        out[tag + "/mean"], out[tag + "/variance"] = d.mean, d.variance
        out[tag + "/entropy"], out[tag + "/log_prob"] = d.entropy(), d.log_prob(v)
        torch.manual_seed(3)
        out[tag + "/sample"] = d.sample()
    return out


@needs_ref
def test_truncated_normal_equals_reference():
    ref = reference_import.load()
    a, b = tn_outputs(ref.TruncatedNormal), tn_outputs(waypoint.TruncatedNormal)
    for k in a:
        torch.testing.assert_close(b[k], a[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")


def test_truncated_normal_golden():
    gold = torch.load(os.path.join(GOLD, "wp_truncnorm.pt"), weights_only=False)
    got = tn_outputs(waypoint.TruncatedNormal)
    for k, v in got.items():
        torch.testing.assert_close(v, gold[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")
    # closed forms against numerical integration of the density
    loc, scale, _ = tn_inputs()
    d = waypoint.TruncatedNormal(loc, scale, 0.25, 4.0)
    xs = torch.linspace(0.25, 4.0, 20001).view(1, -1)
    p = d.log_prob(xs).exp()
    dx = float(xs[0, 1] - xs[0, 0])
    torch.testing.assert_close(torch.trapz(p, dx=dx), torch.ones(64), rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(torch.trapz(p * xs, dx=dx), d.mean.squeeze(1), rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(torch.trapz(-p * d.log_prob(xs), dx=dx), d.entropy().squeeze(1), rtol=1e-3, atol=2e-3)


SMALL = dict(rgb=(12, 6, 6, 3), depth=(12, 8, 8, 1), length=5)


def fill(storage, T, N, seed=5):
    g = torch.Generator().manual_seed(seed)
    for k, v in storage.observations.items():
        v[0].copy_(torch.rand(v[0].shape, generator=g))
    for t in range(T):
        obs = {k: torch.rand(v[0].shape, generator=g) for k, v in storage.observations.items()}
        act = {"pano": torch.randint(0, 13, (N, 1), generator=g), "offset": torch.rand(N, 1, generator=g) - 0.5,
               "distance": torch.rand(N, 1, generator=g) * 3 + 0.25}
        storage.insert(obs, torch.randn(N, 2, 16, generator=g), act, torch.randn(N, 1, generator=g),
                       torch.randn(N, 1, generator=g), torch.randn(N, 1, generator=g),
                       (torch.rand(N, 1, generator=g) > 0.3).to(torch.uint8))
    return torch.randn(N, 1, generator=g)


def storage_outputs(cls, T=5, N=4):
    out = {}
    st = cls(T, N, waypoint.storage_obs_space(**SMALL), 16, num_recurrent_layers=2)
    nv = fill(st, T, N)
    for tag, gae in (("gae", True), ("mc", False)):
        st.compute_returns(nv, gae, 0.99, 0.95)
        out["returns/" + tag] = st.returns.clone()
    adv = st.returns[:-1] - st.value_preds[:-1]
    torch.manual_seed(11)
    for i, sample in enumerate(st.recurrent_generator(adv, 2)):
        for j, part in enumerate(sample):
            for k, v in (part.items() if isinstance(part, dict) else [("", part)]):
                out[f"mb{i}/{j}/{k}"] = v.clone().float()
    st.after_update()
    out["after/masks0"] = st.masks[0].clone().float()
    out["after/prev_pano0"] = st.prev_actions["pano"][0].clone().float()
    out["after/rgb0"] = st.observations["rgb"][0].clone()
    return out


@needs_ref
def test_rollout_storage_equals_reference():
    ref = reference_import.load()
    a, b = storage_outputs(ref.ActionDictRolloutStorage), storage_outputs(waypoint.ActionDictRolloutStorage)
    assert set(a) == set(b)
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


def test_rollout_storage_golden():
    gold = torch.load(os.path.join(GOLD, "wp_storage.pt"), weights_only=False)
    got = storage_outputs(waypoint.ActionDictRolloutStorage)
    assert set(got) == set(gold)
    for k, v in got.items():
        d = torch.tensor([float(v.double().sum()), float(v.double().abs().sum()), float(v.numel())])
        torch.testing.assert_close(d, gold[k], rtol=1e-6, atol=1e-6, msg=lambda m: f"{k}: {m}")


@needs_ref
def test_wddppo_update_equals_reference():
    """The whole ``WDDPPO.update`` (2 mini-batches: evaluate_actions, loss, backward, clip, Adam) on a tiny
    rollout: updated parameters and the six returned scalars match the restated loop."""
    from oracle import cases

    ref = reference_import.load()
    case = dict(cases.CASES["wp_eval"])
    T, N = 2, 2
    res = []
    for impl in ("ref", "oracle"):
        ns = ref if impl == "ref" else waypoint
        pol = cases.build_policy(ns, case)
        cases.set_mode(pol, case)
        st = ns.ActionDictRolloutStorage(T, N, waypoint.storage_obs_space(length=12), 256, num_recurrent_layers=2)
        g = torch.Generator().manual_seed(9)
        from vlnce_b200 import synth

        def rows():
            o = synth.make_waypoint_rows(g, N, length=12)
            return o

        first = rows()
        for k in st.observations:
            st.observations[k][0].copy_(first[k])
        for t in range(T):
            a = synth.make_waypoint_actions(g, N)
            st.insert(rows(), torch.randn(N, 2, 256, generator=g) * 0.3, a, -2 + 0.2 * torch.randn(N, 1, generator=g),
                      torch.randn(N, 1, generator=g) * 0.3, torch.randn(N, 1, generator=g),
                      torch.ones(N, 1, dtype=torch.uint8))
        st.compute_returns(torch.randn(N, 1, generator=g) * 0.3, True, 0.99, 0.95)
        kw = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, offset_regularize_coef=0.1146,
                  pano_entropy_coef=1.0, offset_entropy_coef=1.0, distance_entropy_coef=1.0)
        torch.manual_seed(21)
        if impl == "ref":
            agent = ref.WDDPPO(actor_critic=pol, ppo_epoch=1, num_mini_batch=2, lr=2e-4, eps=1e-5, max_grad_norm=0.2,
                               use_normalized_advantage=False, use_clipped_value_loss=True, **kw)
            scalars = agent.update(st)
        else:
            opt = torch.optim.Adam([p for p in pol.parameters() if p.requires_grad], lr=2e-4, eps=1e-5)
            scalars = waypoint.ppo_update(pol, opt, st, ppo_epoch=1, num_mini_batch=2, max_grad_norm=0.2,
                                          use_clipped_value_loss=True, **kw)
        res.append((scalars, {k: v.detach().clone() for k, v in pol.named_parameters() if v.requires_grad}))
    (sa, pa), (sb, pb) = res
    torch.testing.assert_close(torch.tensor(sa), torch.tensor(sb), rtol=1e-5, atol=1e-6)
    for k in pa:
        torch.testing.assert_close(pb[k], pa[k], rtol=1e-5, atol=1e-6, msg=lambda m: f"{k}: {m}")


if __name__ == "__main__":  # freeze the fixtures from the UNMODIFIED reference (build container only)
    ref = reference_import.load()
    torch.save({k: v.clone() for k, v in tn_outputs(ref.TruncatedNormal).items()}, os.path.join(GOLD, "wp_truncnorm.pt"))
    so = storage_outputs(ref.ActionDictRolloutStorage)
    torch.save({k: torch.tensor([float(v.double().sum()), float(v.double().abs().sum()), float(v.numel())])
                for k, v in so.items()}, os.path.join(GOLD, "wp_storage.pt"))
    print("frozen", len(so), "storage tensors")


def test_select_history_frames_equals_reference_loop():
    """``ddppo_waypoint_trainer.py:190-199`` restated as the reference writes it (per-env loop) vs the batched gather."""
    from vlnce_b200.ppo_step import select_history_frames

    g = torch.Generator().manual_seed(2)
    obs = {"rgb": torch.rand(5, 12, 4, 4, 3, generator=g), "depth": torch.rand(5, 12, 6, 6, 1, generator=g)}
    pano = torch.tensor([[3], [12], [0], [11], [12]])
    want = {"rgb": torch.zeros_like(obs["rgb"][:, 0]), "depth": torch.zeros_like(obs["depth"][:, 0])}
    for i in range(5):
        if int(pano[i]) != 12:                      # actions[i]["action"] != "STOP"
            want["rgb"][i] = obs["rgb"][i, int(pano[i])]
            want["depth"][i] = obs["depth"][i, int(pano[i])]
    got = select_history_frames(obs, pano, 12)
    assert torch.equal(got["rgb"], want["rgb"]) and torch.equal(got["depth"], want["depth"])
