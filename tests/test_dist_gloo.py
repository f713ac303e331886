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
