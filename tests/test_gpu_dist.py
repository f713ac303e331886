"""2-rank NCCL check (needs >= 2 GPUs; skipped otherwise): the all-reduced, averaged gradient
of two trajectory shards equals the single-GPU gradient of the whole batch, and
both ranks hold identical parameters after the fused Adam step."""
import os
import socket
import warnings

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore")


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import vlnce_b200
    from oracle import cases
    from vlnce_b200 import dist as vdist
    from vlnce_b200.il_step import ILTrainerStep

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = f"cuda:{rank}"
    case = dict(cases.CASES["cma_train_features_nopm"], n=4)
    pol = cases.build_policy(vlnce_b200, case).to(dev).eval()
    pol.net.precision = "fp32"
    full = cases.inputs(case)
    shard = tuple(cases._to(v, dev) for v in vdist.shard_batch(full, rank, world))
    step = ILTrainerStep(pol, lr=1e-3)
    step.update_agent(*shard)
    torch.cuda.synchronize()
    g = (step.fp.grad / world).cpu()
    p = step.fp.flat.detach().cpu()
    if rank == 0:  # single-GPU gradient of the concatenated batch, fresh policy
        dist.barrier()
        pol1 = cases.build_policy(vlnce_b200, case).to(dev).eval()
        pol1.net.precision = "fp32"
        s1 = ILTrainerStep(pol1, lr=1e-3)
        s1.world = 1
        s1.update_agent(*(cases._to(v, dev) for v in full))
        torch.cuda.synchronize()
        q.put((rank, g, p, s1.fp.grad.cpu(), s1.fp.flat.detach().cpu()))
    else:
        dist.barrier()
        q.put((rank, g, p, None, None))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_allreduced_gradient_equals_single_gpu_gradient():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r = q.get(timeout=600)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
    g0, p0, g_single, p_single = res[0]
    g1, p1, _, _ = res[1]
    assert torch.equal(g0, g1) and torch.equal(p0, p1), "ranks diverged after the all-reduce"
    # equal-sized shards: mean of per-shard action losses == whole-batch loss (base_il_trainer.py:165)
    torch.testing.assert_close(g0, g_single, rtol=1e-4, atol=1e-7)
    # first Adam step = lr*g/(|g|+eps): entries with |g| ~ eps amplify 1e-9-level gradient noise
    torch.testing.assert_close(p0, p_single, rtol=1e-5, atol=5e-5)


# ---------------------------------------------------------------------------------------------
# DD-PPO step (BASELINE config C4): the same data-parallel identity on the real kernels over NCCL
# ---------------------------------------------------------------------------------------------
def _ppo_worker(rank, world, port, q):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    import torch.distributed as dist
    import vlnce_b200
    from oracle import cases
    from test_dist_gloo import _shard_sample
    from vlnce_b200.ppo_step import WDDPPOStep

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    dev = f"cuda:{rank}"
    case = cases.CASES["wp_eval"]
    t, n = case["t"], case["n"]

    def grad_of(sample, w):
        pol = cases.build_policy(vlnce_b200, case).to(dev)
        pol.net.precision = "fp32"
        cases.set_mode(pol, case)
        agent = WDDPPOStep(pol, max_grad_norm=0.2, offset_entropy_coef=1.0, distance_entropy_coef=1.0)
        agent.world = w
        # equal-length instructions remove quirk W1's batch coupling (see tests/test_dist_gloo.py)
        instr = sample[0]["instruction"]
        sample[0]["instruction"] = instr.view(t, n, -1)[:, :1].expand(t, n, instr.shape[1]).reshape(t * n, -1).contiguous()
        part = _shard_sample(sample, rank, w, t, n) if w > 1 else sample
        part = tuple(cases._to(v, dev) for v in part)
        losses = agent.update_minibatch(part).clone()
        torch.cuda.synchronize()
        return agent.fp.grad.cpu(), losses.cpu(), agent.fp.flat.detach().cpu()

    g, losses, p = grad_of(cases.inputs(case), world)
    dist.barrier()
    single = grad_of(cases.inputs(case), 1) if rank == 0 else (None, None, None)
    q.put((rank, g, losses, p) + single)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_waypoint_ppo_nccl_world2_equals_single_gpu():
    """``WDDPPOStep.update_minibatch`` on two ranks (one env column each, NCCL all-reduce of the flat gradient, the
    1/world average folded into ``vlnce_clip_grad_norm``) == the single-GPU step on the whole mini-batch
    (``ddppo_alg.py:125-131`` with the DD-PPO reducer's averaging, ``ddppo_waypoint_trainer.py:370``)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ppo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r = q.get(timeout=900)
        res[r[0]] = r[1:]
    for p in procs:
        p.join(timeout=60)
    g0, l0, p0, g_single, l_single, p_single = res[0]
    g1, l1, p1 = res[1][:3]
    assert torch.equal(g0, g1) and torch.equal(p0, p1), "ranks diverged after the all-reduce"
    scale = float(g_single.abs().max())
    assert float((g0 - g_single).abs().max()) < 1e-4 * scale
    mean_losses = (l0 + l1) / 2
    torch.testing.assert_close(mean_losses[1:5], l_single[1:5], rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(p0, p_single, rtol=1e-5, atol=5e-5)
