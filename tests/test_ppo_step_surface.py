"""``WDDPPOStep`` keeps the surface the reference's DD-PPO trainer uses on its agent
(``ddppo_waypoint_trainer.py:141-143,425,484-485``; habitat_baselines v0.1.7 ``PPO``): an ``nn.Module`` whose
``state_dict`` keys are prefixed ``actor_critic.``, an ``optimizer`` that checkpoints and accepts an lr scheduler, a
``clip_param`` that can be reassigned between updates.  Host-only (no kernels run)."""
import torch

import vlnce_b200
from oracle import cases
from vlnce_b200.ppo_step import WDDPPOStep


def _agent():
    pol = cases.build_policy(vlnce_b200, cases.CASES["wp_eval"])
    return pol, WDDPPOStep(pol, lr=2.0e-4, clip_param=0.2)


def test_state_dict_keys_and_round_trip():
    pol, agent = _agent()
    assert isinstance(agent, torch.nn.Module)
    sd = agent.state_dict()
    assert sd and all(k.startswith("actor_critic.") for k in sd)
    assert {k[len("actor_critic."):] for k in sd} == set(pol.state_dict())
    pol2, agent2 = _agent()
    with torch.no_grad():
        for p in pol2.parameters():
            p.add_(1.0)
    agent2.load_state_dict(sd)                      # ddppo_waypoint_trainer.py:143
    assert all(torch.equal(a, b) for a, b in zip(pol.state_dict().values(), pol2.state_dict().values()))


def test_optimizer_checkpoints_and_follows_a_scheduler():
    _, agent = _agent()
    opt = agent.optimizer
    assert isinstance(opt, torch.optim.Optimizer) and opt.param_groups[0]["lr"] == 2.0e-4
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lambda e: 0.5 ** e)   # :425 linear_decay analogue
    agent._prepare()
    agent.fp.exp_avg.fill_(0.25)
    agent.fp.exp_avg_sq.fill_(0.5)
    agent.step_id = 7
    opt._step_count = 1  # silence the "scheduler before optimizer" warning: the fused step is not optimizer.step()
    sched.step()
    assert abs(agent.lr - 1.0e-4) < 1e-12            # the fused Adam launch reads param_groups[0]["lr"]
    saved = opt.state_dict()
    _, other = _agent()
    other.optimizer.load_state_dict(saved)
    assert other.step_id == 7 and float(other.fp.exp_avg[3]) == 0.25 and float(other.fp.exp_avg_sq[5]) == 0.5
    assert abs(other.lr - 1.0e-4) < 1e-12


def test_clip_param_is_read_at_update_time():
    _, agent = _agent()
    agent.clip_param = 0.05                           # use_linear_clip_decay (ddppo_waypoint_trainer.py:517-520)
    assert "ppo_consts[0] = float(self.clip_param)" in open(vlnce_b200.ppo_step.__file__).read()
    assert agent.ppo_consts[0] != 0.05                # refreshed inside update_minibatch, not frozen at construction
