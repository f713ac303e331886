"""Named parity cases shared by the golden-vector generator and the tests.
TEST INFRASTRUCTURE (see oracle/__init__).

A case is a pure function of seeds: model config, weight seed, input seed, shapes.
``run_case(impl, name)`` runs it through an implementation namespace (the imported
reference, or ``oracle.vlnce``) and returns a flat dict of small tensors / floats —
what ``tests/golden/*.pt`` stores.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from vlnce_b200 import synth

from . import hb017, vlnce

CASES = {
    # BASELINE config C2 shape family (CMA + bi-LSTM + progress monitor), small batch
    "cma_act": dict(policy="CMAPolicy", pm=True, kind="act", n=3, hw=256, length=80,
                    wseed=11, iseed=101, bn_train=False),
    "cma_act_bntrain": dict(policy="CMAPolicy", pm=True, kind="act", n=3, hw=256, length=80,
                            wseed=11, iseed=101, bn_train=True),
    "cma_act_rgb224": dict(policy="CMAPolicy", pm=False, kind="act", n=2, hw=256, rgb_hw=224,
                           length=24, wseed=12, iseed=102, bn_train=False),
    "cma_train_pixels": dict(policy="CMAPolicy", pm=True, kind="train", t=3, n=2, hw=256,
                             length=80, wseed=13, iseed=103, bn_train=False, features=False),
    "cma_train_features": dict(policy="CMAPolicy", pm=True, kind="train", t=5, n=4, hw=256,
                               length=40, wseed=14, iseed=104, bn_train=False, features=True),
    "cma_train_features_nopm": dict(policy="CMAPolicy", pm=False, kind="train", t=4, n=3, hw=256,
                                    length=16, wseed=15, iseed=105, bn_train=False,
                                    features=True),
    # BASELINE config C1: Seq2Seq, batch 2
    "seq2seq_act": dict(policy="Seq2SeqPolicy", pm=False, kind="act", n=2, hw=256, length=80,
                        wseed=16, iseed=106, bn_train=False),
    # Seq2Seq training (seq2seq_r2n.yaml recipe family): raw pixels with the progress monitor, and the
    # DAgger cached-feature layout with the prev-action embedding (seq2seq_policy.py:158-162)
    "seq2seq_train_pixels": dict(policy="Seq2SeqPolicy", pm=True, kind="train", t=3, n=2, hw=256,
                                 length=40, wseed=17, iseed=107, bn_train=False, features=False),
    "seq2seq_train_features": dict(policy="Seq2SeqPolicy", pm=False, kind="train", t=4, n=3, hw=256,
                                   length=24, wseed=18, iseed=108, bn_train=False, features=True,
                                   prev_action=True),
}


def model_config(case) -> hb017.Config:
    over = dict(PROGRESS_MONITOR=dict(use=bool(case["pm"]), alpha=1.0))
    if case["policy"] == "CMAPolicy":
        return vlnce.cma_model_config(**over)
    if case.get("prev_action"):
        over["SEQ2SEQ"] = dict(use_prev_action=True)
    return vlnce.default_model_config(policy_name="Seq2SeqPolicy", **over)


def obs_space(case):
    rgb_hw = case.get("rgb_hw", case["hw"])
    return hb017.Dict({"rgb": hb017.Box(0, 255, (rgb_hw, rgb_hw, 3), "uint8"),
                       "depth": hb017.Box(0.0, 1.0, (case["hw"], case["hw"], 1), "float32")})


def build_policy(impl, case, num_actions=4):
    """impl: namespace with CMAPolicy / Seq2SeqPolicy (reference import, oracle.vlnce
    or the CUDA package)."""
    cfg = hb017.Config(MODEL=model_config(case), TORCH_GPU_ID=0)
    cls = getattr(impl, case["policy"])
    pol = cls.from_config(cfg, obs_space(case), hb017.Discrete(num_actions))
    pol.load_state_dict(synth.seeded_state_dict(pol, case["wseed"]))
    return pol


def set_mode(pol, case):
    """``bn_train=False``: whole policy in eval() (what eval/inference and DD-PPO use,
    ``base_il_trainer.py:278``); True: the constructor default."""
    pol.train(bool(case["bn_train"]))


def layers_of(case):
    return 2 if case["policy"] == "CMAPolicy" else 1


def inputs(case):
    if case["kind"] == "act":
        return synth.make_act_batch(case["iseed"], case["n"], hw=case["hw"],
                                    rgb_hw=case.get("rgb_hw"), length=case["length"],
                                    layers=layers_of(case))
    return synth.make_train_batch(case["iseed"], case["t"], case["n"], hw=case["hw"],
                                  rgb_hw=case.get("rgb_hw"), length=case["length"],
                                  features=case["features"],
                                  rgb_feature_hw=1 if case["policy"] == "Seq2SeqPolicy" else 4)


def grad_digest(g: torch.Tensor) -> torch.Tensor:
    g = g.detach().double().reshape(-1)
    head = torch.zeros(8, dtype=torch.float64)
    head[: min(8, g.numel())] = g[:8]
    return torch.cat([torch.stack([g.sum(), g.abs().sum(), g.norm()]), head]).float()


def _to(x, device):
    if device is None:
        return x
    if isinstance(x, dict):
        return {k: v.to(device) for k, v in x.items()}
    return x.to(device)


def run_case(impl, name, aux=None, device=None, precision=None):
    """Runs ``name`` through ``impl``; ``aux`` = that implementation's AuxLosses
    singleton.  ``device`` / ``precision`` select the CUDA package's placement and trunk
    arithmetic.  Returns {key: cpu tensor}."""
    case = CASES[name]
    aux = aux if aux is not None else impl.AuxLosses
    torch.manual_seed(0)
    pol = build_policy(impl, case)
    if device is not None:
        pol = pol.to(device)
    if precision is not None:
        pol.net.precision = precision
    set_mode(pol, case)
    out = {}
    if case["kind"] == "act":
        obs, h, prev, masks = (_to(v, device) for v in inputs(case))
        aux.deactivate()
        with torch.no_grad():
            dist = pol.build_distribution(obs, h, prev, masks)
            action, h_out = pol.act(obs, h, prev, masks, deterministic=True)
        out["logits"] = dist.logits.detach().cpu().clone()
        out["action"] = action.cpu().clone()
        out["rnn_states_out"] = h_out.detach().cpu().clone()
        return out
    obs, prev, masks, actions, weights = (_to(v, device) for v in inputs(case))
    t, n = actions.shape
    aux.activate()
    aux.clear()
    h0 = torch.zeros(n, pol.net.num_recurrent_layers, 512, device=device)
    dist = pol.build_distribution(obs, h0, prev, masks)
    logits = dist.logits.view(t, n, -1)
    ce = F.cross_entropy(logits.permute(0, 2, 1), actions, reduction="none")
    action_loss = ((weights * ce).sum(0) / weights.sum(0)).mean()
    aux_loss = aux.reduce((weights > 0).view(-1))
    loss = action_loss + aux_loss
    for p in pol.parameters():
        p.grad = None
    loss.backward()
    aux.deactivate()
    out["logits"] = dist.logits.detach().cpu().clone()
    out["loss"] = torch.tensor(float(loss.detach()))
    out["action_loss"] = torch.tensor(float(action_loss.detach()))
    out["aux_loss"] = torch.tensor(float(aux_loss.detach() if torch.is_tensor(aux_loss) else aux_loss))
    for k, p in pol.named_parameters():
        if p.requires_grad:
            g = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu()
            out["grad/" + k] = g.clone() if g.numel() <= 4096 else grad_digest(g)
    return out
