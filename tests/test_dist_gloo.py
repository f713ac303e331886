"""World-size-2 ``gloo`` test (CPU) of the data-parallel host logic: sharding of the
collate_fn layout and the single flat-gradient all-reduce inside ``ILTrainerStep``."""
import os
import socket
import warnings

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

warnings.filterwarnings("ignore")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import vlnce_b200
    from oracle import cases
    from vlnce_b200 import backward, dist as vdist, il_step, lib

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    case = dict(cases.CASES["cma_train_features_nopm"], n=4)
    pol = cases.build_policy(vlnce_b200, case).eval()
    full = cases.inputs(case)
    shard = vdist.shard_batch(full, rank, world)
    t, n = full[3].shape
    assert shard[3].shape == (t, n // world) and shard[0]["instruction"].shape[0] == t * n // world
    # rows of trajectory j on this rank are rows (t, rank*per + j) of the full batch
    per = n // world
    want = full[0]["rgb_features"].view(t, n, -1)[:, rank * per:(rank + 1) * per].reshape(t * per, -1)
    assert torch.equal(shard[0]["rgb_features"].reshape(t * per, -1), want)

    orig = backward.cma_backward

    def fake_backward(eng, head, g, d_raw, d_x, d_hat):  # stands in for the CUDA kernels on CPU
        orig(eng, head, g, d_raw, d_x, d_hat)
        for v in g.values():
            v.fill_(float(rank + 1))

    backward.cma_backward = fake_backward  # CMAEngine.run_backward resolves it at call time
    step = il_step.ILTrainerStep(pol, lr=1e-3)
    assert step.world == world
    with lib.dry_run() as rec:
        step.update_agent(*shard, return_tensor=True)
    assert rec.calls[-1] == "vlnce_adam_step"
    expect = float(sum(r + 1 for r in range(world)))  # every entry: sum over ranks of (rank+1)
    ok = float(step.fp.grad.min()) == expect and float(step.fp.grad.max()) == expect
    out.put((rank, ok, step.fp.total))
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and res[0][2] == res[1][2]


# ---------------------------------------------------------------------------------------------
# waypoint / DD-PPO path: a NUMERICAL data-parallel check on the CPU — the launch plan of each rank runs on
# host memory through the numpy model of the kernels (tests/abi_emulator.py), the all-reduce is real (gloo)
# ---------------------------------------------------------------------------------------------
def _shard_sample(sample, rank, world, t, n):
    from vlnce_b200 import dist as vdist

    per = n // world

    def rows(x):
        return vdist.shard_rows(x, t, n, rank, world)

    obs, h0, actions, prev, v_old, ret, masks, lp_old, adv = sample
    return ({k: rows(v) for k, v in obs.items()}, h0[rank * per:(rank + 1) * per].contiguous(),
            {k: rows(v) for k, v in actions.items()}, {k: rows(v) for k, v in prev.items()}, rows(v_old), rows(ret),
            rows(masks), rows(lp_old), rows(adv))


def _ppo_grad(world, rank, sample, case):
    """Clipped, averaged flat gradient + the loss scalars after one fused PPO mini-batch on this rank's shard."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import vlnce_b200
    from abi_emulator import Emulator
    from oracle import cases
    from test_host_emulated import substitute_trunks
    from vlnce_b200 import lib
    from vlnce_b200.ppo_step import WDDPPOStep

    pol = cases.build_policy(vlnce_b200, case)
    cases.set_mode(pol, case)
    substitute_trunks(pol, case)
    agent = WDDPPOStep(pol, max_grad_norm=0.2, offset_entropy_coef=1.0, distance_entropy_coef=1.0)
    assert agent.world == world
    t, n = case["t"], case["n"]
    # Quirk W1 makes the instruction attention depend on the LONGEST instruction of the rows processed together
    # (the padded length), in the reference as well: sharding a ragged batch changes it.  Equal-length rows
    # remove that batch coupling so that the data-parallel identity is exact.
    instr = sample[0]["instruction"]
    sample[0]["instruction"] = instr.view(t, n, -1)[:, :1].expand(t, n, instr.shape[1]).reshape(t * n, -1).contiguous()
    shard = _shard_sample(sample, rank, world, t, n) if world > 1 else sample
    with lib.dry_run(executor=Emulator()) as rec:
        losses = agent.update_minibatch(shard).clone()
    assert rec.calls[-1] == "vlnce_adam_step" and rec.calls.count("vlnce_clip_grad_norm") == 1
    return agent.fp.grad.clone(), losses, float(agent._clip_ws[1024])


def _ppo_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import cases

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    case = cases.CASES["wp_eval"]
    grad, losses, norm = _ppo_grad(world, rank, cases.inputs(case), case)
    out.put((rank, grad.numpy(), losses.numpy(), norm))
    dist.destroy_process_group()


def test_waypoint_ppo_world2_equals_single_process():
    """Two ranks, one env column each, ONE all-reduce of the flat gradient, the 1/world average folded into
    clip_grad_norm_: the clipped gradient equals the single-process gradient of the whole mini-batch
    (mean losses over equal shards), as with the reference's DD-PPO reducer (ddppo_waypoint_trainer.py:370)."""
    import numpy as np
    from oracle import cases

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ppo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    case = cases.CASES["wp_eval"]
    want_grad, want_losses, want_norm = _ppo_grad(1, 0, cases.inputs(case), case)
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g0, g1 = res[0][1], res[1][1]
    assert np.array_equal(g0, g1)                      # replicas stay in lock-step
    scale = float(want_grad.abs().max())
    assert float(np.abs(g0 - want_grad.numpy()).max()) < 2e-5 * scale
    assert abs(res[0][3] - want_norm) < 1e-4 * want_norm
    mean_losses = (res[0][2] + res[1][2]) / 2          # per-rank means over equal shards
    assert np.allclose(mean_losses[1:5], want_losses.numpy()[1:5], rtol=1e-5, atol=1e-7)


# ---------------------------------------------------------------------------------------------
# imitation-learning step (the benchmarked path): the same NUMERICAL data-parallel identity
# ---------------------------------------------------------------------------------------------
def _il_grad(world, rank):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    import vlnce_b200
    from abi_emulator import Emulator
    from oracle import cases
    from vlnce_b200 import dist as vdist, il_step, lib

    case = dict(cases.CASES["cma_train_features_nopm"], n=4)
    pol = cases.build_policy(vlnce_b200, case).eval()
    full = cases.inputs(case)
    batch = vdist.shard_batch(full, rank, world) if world > 1 else full
    step = il_step.ILTrainerStep(pol, lr=1e-3)
    assert step.world == world
    before = None
    with lib.dry_run(executor=Emulator()) as rec:
        step.compute_grads(*batch)
        local = step.fp.grad.clone()
        before = step.fp.flat.clone()
        step.apply_grads(True)
    assert rec.calls[-1] == "vlnce_adam_step"
    return step.fp.grad.clone() / world, local, step.fp.flat.clone() - before, step._losses.clone()


def _il_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    avg, local, delta, losses = _il_grad(world, rank)
    out.put((rank, avg.numpy(), local.numpy(), delta.numpy(), losses.numpy()))
    dist.destroy_process_group()


def test_il_step_world2_equals_single_process():
    """ILTrainerStep on two ranks (2 of 4 trajectories each): the all-reduced gradient with the 1/world factor folded
    into Adam equals the single-process gradient of the whole batch (per-trajectory weighted means, equal shards,
    ``base_il_trainer.py:159-165``), and both replicas apply the same parameter update."""
    import numpy as np

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_il_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    want_grad, _, want_delta, want_losses = _il_grad(1, 0)
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][3], res[1][3])   # replicas in lock-step
    assert not np.array_equal(res[0][2], res[1][2])                                       # ...from different shards
    scale = float(want_grad.abs().max())
    assert float(np.abs(res[0][1] - want_grad.numpy()).max()) < 2e-5 * scale
    assert float(np.abs(res[0][3] - want_delta.numpy()).max()) < 2e-5        # same Adam update (lr 1e-3)
    assert abs((res[0][4][1] + res[1][4][1]) / 2 - float(want_losses[1])) < 1e-5 * abs(float(want_losses[1]))
