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
    # BASELINE config C5 shape family (rxr_cma_en.yaml): precomputed multilingual-BERT instruction features
    # (768-d, zero rows beyond the instruction), 6 actions; L cut from 512 to keep the CPU oracle fast
    "cma_rxr_act": dict(policy="CMAPolicy", pm=False, kind="act", n=3, hw=256, rgb_hw=224, length=48, wseed=19,
                        iseed=109, bn_train=False, rxr=768, actions=6),
    "cma_rxr_train_features": dict(policy="CMAPolicy", pm=True, kind="train", t=3, n=3, hw=256, length=40, wseed=20,
                                   iseed=110, bn_train=False, features=True, rxr=768, actions=6, coef=1.9),
    # modality ablations (config/default.py:217-219; the paper's ablation studies)
    "cma_train_features_abl_i": dict(policy="CMAPolicy", pm=True, kind="train", t=3, n=3, hw=256, length=24, wseed=31,
                                     iseed=131, bn_train=False, features=True, ablate=("instruction",)),
    "cma_train_features_abl_v": dict(policy="CMAPolicy", pm=False, kind="train", t=3, n=3, hw=256, length=24, wseed=32,
                                     iseed=132, bn_train=False, features=True, ablate=("rgb", "depth")),
    "seq2seq_train_features_abl": dict(policy="Seq2SeqPolicy", pm=False, kind="train", t=3, n=3, hw=256, length=24,
                                       wseed=33, iseed=133, bn_train=False, features=True, prev_action=True,
                                       ablate=("instruction", "rgb")),
    # STATE_ENCODER.rnn_type = "LSTM" (config/default.py:259 allows GRU | LSTM; rnn_states carry (h, c) per encoder)
    "cma_act_lstm": dict(policy="CMAPolicy", pm=True, kind="act", n=3, hw=256, length=24, wseed=35, iseed=135,
                         bn_train=False, state_rnn="LSTM"),
    "cma_train_features_lstm": dict(policy="CMAPolicy", pm=True, kind="train", t=4, n=3, hw=256, length=24, wseed=36,
                                    iseed=136, bn_train=False, features=True, state_rnn="LSTM"),
    "seq2seq_train_features_lstm": dict(policy="Seq2SeqPolicy", pm=True, kind="train", t=4, n=3, hw=256, length=24,
                                        wseed=37, iseed=137, bn_train=False, features=True, prev_action=True,
                                        state_rnn="LSTM"),
    "wp_eval_lstm": dict(policy="WaypointPolicy", kind="wp_eval", t=3, n=2, length=24, wseed=38, iseed=138,
                         state_rnn="LSTM"),
    # INSTRUCTION_ENCODER.rnn_type = "GRU" (instruction_encoder.py:27): bi-GRU per-token outputs (CMA, waypoint),
    # uni-directional final state (Seq2Seq)
    "cma_train_features_igru": dict(policy="CMAPolicy", pm=True, kind="train", t=3, n=3, hw=256, length=24, wseed=39,
                                    iseed=139, bn_train=False, features=True, instr_rnn="GRU"),
    "seq2seq_train_features_igru": dict(policy="Seq2SeqPolicy", pm=False, kind="train", t=3, n=3, hw=256, length=24,
                                        wseed=40, iseed=140, bn_train=False, features=True, instr_rnn="GRU"),
    "wp_eval_igru": dict(policy="WaypointPolicy", kind="wp_eval", t=2, n=2, length=24, wseed=41, iseed=141,
                         instr_rnn="GRU", state_rnn="LSTM"),
    # MODEL.normalize_rgb = True (config/default.py:216; resnet_encoders.py:171-192): ImageNet mean / std on the RGB
    "cma_act_nrgb": dict(policy="CMAPolicy", pm=False, kind="act", n=2, hw=256, rgb_hw=224, length=24, wseed=42,
                         iseed=142, bn_train=False, normalize_rgb=True),
    "wp_act_nrgb": dict(policy="WaypointPolicy", kind="wp_act", n=2, length=24, wseed=43, iseed=143,
                        normalize_rgb=True, rgb_u8=True),
    # trainable conv trunks (config/default.py:242,253; next-row N4): gradients through both ResNets, train-mode
    # BatchNorm in the RGB trunk (``cnn.train(trainable)`` + ``policy.train()``), GroupNorm in the depth trunk
    "cma_train_pixels_trunks": dict(policy="CMAPolicy", pm=True, kind="train", t=2, n=2, hw=256, rgb_hw=224, length=16,
                                    wseed=44, iseed=144, bn_train=True, features=False, trainable=("rgb", "depth")),
    "seq2seq_train_pixels_trunks": dict(policy="Seq2SeqPolicy", pm=False, kind="train", t=2, n=2, hw=256, length=16,
                                        wseed=45, iseed=145, bn_train=False, features=False,
                                        trainable=("rgb", "depth")),
    "wp_eval_abl": dict(policy="WaypointPolicy", kind="wp_eval", t=2, n=2, length=24, wseed=34, iseed=134,
                        ablate=("instruction", "depth")),
    # trained-like action heads (SURVEY hard-part 2): the head's logit spread is far above every tolerance, so the
    # argmax action is asserted bit-exact on EVERY row, 16-bit trunks included (``head_scale`` multiplies the seeded
    # head weight; a trained CMA head separates its top-2 logits by O(1))
    "cma_act_sharp": dict(policy="CMAPolicy", pm=True, kind="act", n=6, hw=256, length=80, wseed=61, iseed=161,
                          bn_train=False, head_scale=24.0),
    "seq2seq_act_sharp": dict(policy="Seq2SeqPolicy", pm=False, kind="act", n=4, hw=256, length=80, wseed=62,
                              iseed=162, bn_train=False, head_scale=24.0),
    # BASELINE config C5 at its real instruction shape: (R, 512, 768) multilingual-BERT features
    # (habitat_extensions/sensors.py:176-196, rxr_cma_en.yaml:43-48), 6 actions, 224 / 256 frames
    "cma_rxr_act_l512": dict(policy="CMAPolicy", pm=False, kind="act", n=2, hw=256, rgb_hw=224, length=512, wseed=63,
                             iseed=163, bn_train=False, rxr=768, actions=6, head_scale=24.0),
    "cma_rxr_train_features_l512": dict(policy="CMAPolicy", pm=True, kind="train", t=2, n=2, hw=256, length=512,
                                        wseed=64, iseed=164, bn_train=False, features=True, rxr=768, actions=6,
                                        coef=1.9),
    # "+ waypoint head" of C5: WaypointPolicy fed the RxR features (the net still asserts an ``instruction`` key,
    # waypoint_predictors.py:310-315; its encoder reads ``rxr_instruction``, instruction_encoder.py:74-78)
    "wp_act_rxr": dict(policy="WaypointPolicy", kind="wp_act", n=2, length=512, wseed=65, iseed=165, rxr=768),
    # Waypoint path (BASELINE config C4 dims, 1-wpn-cc.yaml): act on a fresh/continuing env pair,
    # and one PPO mini-batch (evaluate_actions -> WDDPPO losses -> gradients)
    "wp_act": dict(policy="WaypointPolicy", kind="wp_act", n=2, length=40, wseed=21, iseed=201),
    "wp_act_u8": dict(policy="WaypointPolicy", kind="wp_act", n=2, length=24, wseed=22, iseed=202, rgb_u8=True),
    "wp_eval": dict(policy="WaypointPolicy", kind="wp_eval", t=3, n=2, length=32, wseed=23, iseed=203),
    # heading-prediction variants (5-hpn-_c.yaml: fixed 0.25 m distance; 6-hpn-__.yaml: neither head sampled)
    "wp_act_hpn": dict(policy="WaypointPolicy", kind="wp_act", n=2, length=24, wseed=24, iseed=204,
                       waypoint=dict(predict_distance=False)),
    "wp_eval_hpn": dict(policy="WaypointPolicy", kind="wp_eval", t=2, n=2, length=24, wseed=25, iseed=205,
                        waypoint=dict(predict_distance=False, predict_offset=False)),
    # discrete heads: 2-wpn-dc.yaml (discrete distance, continuous offset), 3-wpn-dd.yaml (both discrete),
    # 4-wpn-d_.yaml (discrete distance, offset not predicted)
    "wp_act_dc": dict(policy="WaypointPolicy", kind="wp_act", n=2, length=24, wseed=46, iseed=206,
                      waypoint=dict(continuous_distance=False)),
    "wp_act_dd": dict(policy="WaypointPolicy", kind="wp_act", n=3, length=24, wseed=55, iseed=207,
                      waypoint=dict(continuous_distance=False, continuous_offset=False)),
    "wp_eval_dd": dict(policy="WaypointPolicy", kind="wp_eval", t=2, n=2, length=24, wseed=28, iseed=208,
                       waypoint=dict(continuous_distance=False, continuous_offset=False)),
    "wp_eval_d_": dict(policy="WaypointPolicy", kind="wp_eval", t=2, n=2, length=24, wseed=29, iseed=209,
                       waypoint=dict(continuous_distance=False, predict_offset=False)),
}

WP_OBS_SPACE = dict(rgb=(12, 224, 224, 3), depth=(12, 256, 256, 1))


def oracle_impl(case):
    """The restatement namespace that implements ``case`` (oracle.vlnce or oracle.waypoint)."""
    if case["policy"] == "WaypointPolicy":
        from . import waypoint

        return waypoint
    return vlnce


def model_config(case) -> hb017.Config:
    abl = {"ablate_" + k: True for k in case.get("ablate", ())}
    if case.get("state_rnn"):
        abl["STATE_ENCODER"] = dict(rnn_type=case["state_rnn"])
    for m in case.get("trainable", ()):
        abl.setdefault(m.upper() + "_ENCODER", {})["trainable"] = True
    if case.get("normalize_rgb"):
        abl["normalize_rgb"] = True
    if case.get("instr_rnn"):
        abl["INSTRUCTION_ENCODER"] = dict(rnn_type=case["instr_rnn"])
    if case["policy"] == "WaypointPolicy":
        from . import waypoint

        if case.get("rxr"):
            abl.setdefault("INSTRUCTION_ENCODER", {}).update(sensor_uuid="rxr_instruction",
                                                             embedding_size=int(case["rxr"]))
        return waypoint.waypoint_model_config(WAYPOINT=dict(case.get("waypoint", {})), **abl)
    over = dict(PROGRESS_MONITOR=dict(use=bool(case["pm"]), alpha=1.0), **abl)
    if case.get("rxr"):  # rxr_cma_en.yaml:43-48
        over["INSTRUCTION_ENCODER"] = dict(sensor_uuid="rxr_instruction", embedding_size=int(case["rxr"]))
    if case["policy"] == "CMAPolicy":
        return vlnce.cma_model_config(**over)
    if case.get("prev_action"):
        over["SEQ2SEQ"] = dict(use_prev_action=True)
    return vlnce.default_model_config(policy_name="Seq2SeqPolicy", **over)


def obs_space(case):
    if case["policy"] == "WaypointPolicy":
        return hb017.Dict({"rgb": hb017.Box(0, 255, WP_OBS_SPACE["rgb"], "uint8"),
                           "depth": hb017.Box(0.0, 1.0, WP_OBS_SPACE["depth"], "float32")})
    rgb_hw = case.get("rgb_hw", case["hw"])
    return hb017.Dict({"rgb": hb017.Box(0, 255, (rgb_hw, rgb_hw, 3), "uint8"),
                       "depth": hb017.Box(0.0, 1.0, (case["hw"], case["hw"], 1), "float32")})


class _skip_random_init:
    """Every parameter and buffer is overwritten by the seeded ``state_dict`` right after construction (strict load),
    so the constructors' random initialisers (kaiming / orthogonal QR over ~30 M weights, about half of the build time
    of the ~350 policies the test-suite constructs) are turned into no-ops while the policy is being built."""

    NAMES = ("kaiming_normal_", "kaiming_uniform_", "orthogonal_", "normal_", "uniform_", "xavier_uniform_",
             "xavier_normal_", "trunc_normal_")

    def __enter__(self):
        init = torch.nn.init
        self.saved = {n: getattr(init, n) for n in self.NAMES}
        for n in self.NAMES:
            setattr(init, n, lambda tensor, *a, **k: tensor)

    def __exit__(self, *exc):
        for n, f in self.saved.items():
            setattr(torch.nn.init, n, f)


def build_policy(impl, case, num_actions=None):
    """impl: namespace with CMAPolicy / Seq2SeqPolicy (reference import, oracle.vlnce
    or the CUDA package)."""
    cfg = hb017.Config(MODEL=model_config(case), TORCH_GPU_ID=0,
                       TASK_CONFIG=dict(TASK=dict(PANO_ROTATIONS=12)))
    cls = getattr(impl, case["policy"])
    num_actions = num_actions or case.get("actions", 4)
    with _skip_random_init():
        pol = cls.from_config(cfg, obs_space(case), hb017.Discrete(num_actions))
    sd = synth.seeded_state_dict(pol, case["wseed"])
    if case.get("head_scale"):
        sd["action_distribution.linear.weight"] = sd["action_distribution.linear.weight"] * float(case["head_scale"])
    pol.load_state_dict(sd, strict=True)
    return pol


def set_mode(pol, case):
    if case["policy"] == "WaypointPolicy":
        # ddppo_waypoint_trainer.py:496 (rollouts: agent.eval()) / :526-530 (updates: train() with both
        # visual encoders forced to eval) — the frozen trunks always run with eval-mode norms
        pol.train(case["kind"] == "wp_eval")
        pol.net.rgb_encoder.eval()
        pol.net.depth_encoder.eval()
        return
    """``bn_train=False``: whole policy in eval() (what eval/inference and DD-PPO use,
    ``base_il_trainer.py:278``); True: the constructor default (SURVEY Q1)."""
    pol.train(bool(case["bn_train"]))


def layers_of(case):
    """rnn_states slots: one per GRU state encoder, two (h, c) per LSTM one"""
    encoders = 1 if case["policy"] == "Seq2SeqPolicy" else 2
    return encoders * (2 if case.get("state_rnn") == "LSTM" else 1)


def inputs(case):
    if case["kind"] in ("wp_act", "wp_eval"):
        wp = case.get("waypoint", {})
        classes = dict(distance=0 if wp.get("continuous_distance", True) else 6,
                       offset=0 if wp.get("continuous_offset", True) else 7)
        if case["kind"] == "wp_act":
            batch = synth.make_waypoint_act_batch(case["iseed"], case["n"], length=case["length"],
                                                  rgb_u8=case.get("rgb_u8", False), classes=classes)
            if case.get("rxr"):  # RxR features beside the (still required) token ids
                gen = torch.Generator().manual_seed(case["iseed"] + 7)
                batch[0]["rxr_instruction"] = synth.make_rxr_instruction(gen, case["n"], case["length"], case["rxr"],
                                                                         min_len=case["length"] // 2)
            return batch
        return synth.make_waypoint_eval_batch(case["iseed"], case["t"], case["n"], length=case["length"],
                                              classes=classes, layers=layers_of(case))
    if case["kind"] == "act":
        return synth.make_act_batch(case["iseed"], case["n"], hw=case["hw"],
                                    rgb_hw=case.get("rgb_hw"), length=case["length"],
                                    layers=layers_of(case), num_actions=case.get("actions", 4),
                                    rxr_dim=case.get("rxr", 0))
    return synth.make_train_batch(case["iseed"], case["t"], case["n"], hw=case["hw"],
                                  rgb_hw=case.get("rgb_hw"), length=case["length"],
                                  features=case["features"],
                                  rgb_feature_hw=1 if case["policy"] == "Seq2SeqPolicy" else 4,
                                  num_actions=case.get("actions", 4), coef=case.get("coef", 3.2),
                                  rxr_dim=case.get("rxr", 0))


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
    aux = aux if aux is not None else getattr(impl, "AuxLosses", None)
    torch.manual_seed(0)
    pol = build_policy(impl, case)
    if device is not None:
        pol = pol.to(device)
    if precision is not None:
        pol.net.precision = precision
    set_mode(pol, case)
    out = {}
    if case["kind"].startswith("wp_"):
        return _run_waypoint(impl, pol, case, device)
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


PPO_KW = dict(clip_param=0.2, value_loss_coef=0.5, entropy_coef=0.01, use_clipped_value_loss=True,
              offset_regularize_coef=0.1146, pano_entropy_coef=1.0, offset_entropy_coef=1.0,
              distance_entropy_coef=1.0)  # config/default.py:179-201; offset/distance entropy coefficients
# raised from the recipe's 0.0 so that the TruncatedNormal entropy gradients are exercised


def _run_waypoint(impl, pol, case, device):
    """``wp_act``: deterministic ``WaypointPolicy.act`` + ``get_value``.  ``wp_eval``: one PPO mini-batch
    — ``evaluate_actions`` -> the WDDPPO loss (``common/ddppo_alg.py:56-124``) -> gradients."""
    out = {}
    if case["kind"] == "wp_act":
        obs, h, prev, masks = (_to(v, device) for v in inputs(case))
        with torch.no_grad():
            (value, actions, elems, modes, variances, lp, h_out, dist) = pol.act(
                obs, h, {k: v.clone() for k, v in prev.items()}, masks, deterministic=True)
            v2 = pol.get_value(obs, h, {k: v.clone() for k, v in prev.items()}, masks)
        out["value"], out["get_value"] = value.cpu().clone(), v2.cpu().clone()
        out["pano_stop_logits"] = dist.logits.detach().cpu().clone()
        out["action_log_probs"] = lp.cpu().clone()
        out["rnn_states_out"] = h_out.cpu().clone()
        for k, v in elems.items():
            out["action/" + k] = v.cpu().clone()
        for k, v in modes.items():
            out["mode/" + k] = v.cpu().clone()
        for k, v in variances.items():  # torch's Categorical.variance (discrete heads) is NaN
            out["variance/" + k] = torch.nan_to_num(v.float().cpu().clone(), nan=-777.0)
        out["stop"] = torch.tensor([a["action"] == "STOP" for a in actions])
        out["r_theta"] = torch.tensor([[a["action"]["action_args"]["r"], a["action"]["action_args"]["theta"]]
                                       if a["action"] != "STOP" else [0.0, 0.0] for a in actions])
        return out
    sample = tuple(_to(v, device) for v in inputs(case))
    for p in pol.parameters():
        p.grad = None
    loss, parts = wddppo_loss(impl, pol, sample)
    loss.backward()
    out["loss"] = torch.tensor(float(loss.detach()))
    for k, v in parts.items():
        out["ppo/" + k] = v.detach().cpu().clone() if torch.is_tensor(v) else torch.tensor(float(v))
    for k, p in pol.named_parameters():
        if p.requires_grad:
            g = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().cpu()
            out["grad/" + k] = g.clone() if g.numel() <= 4096 else grad_digest(g)
    return out


def wddppo_loss(impl, pol, sample):
    """The mini-batch loss of ``WDDPPO.update`` through ``impl``'s ``evaluate_actions``.  The arithmetic after
    ``evaluate_actions`` is the restated one (``oracle.waypoint.wddppo_losses``) for every implementation; the
    reference's own ``WDDPPO.update`` is pinned separately in tests/test_oracle_vs_reference.py."""
    from . import waypoint

    loss, parts = waypoint.wddppo_losses(pol, sample, **PPO_KW)
    return loss, parts
