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
