"""SURVEY §8f N4 building blocks: the fp32 trunk-backward entry points (csrc/trunk_bwd.cu) against torch autograd.
``run(dev)`` is shared by the host check (numpy/torch model of the kernels through the same C-ABI calls — it pins the
call conventions and the expectations) and the GPU test in tests/test_zz_gpu_variants.py (the CUDA kernels)."""
import torch
import torch.nn.functional as Fn

from vlnce_b200 import lib


def run(dev, C):
    ptr = lib.ptr
    st = lib.stream() if dev != "cpu" else 0
    g0 = torch.Generator().manual_seed(3)
    rnd = lambda *s: torch.randn(*s, generator=g0)  # noqa: E731
    # ---- conv dgrad / wgrad (NHWC / KRSC) over the trunk's geometries
    for (n, h, w, c, k, r, stride, pad) in [(2, 9, 8, 6, 5, 3, 1, 1), (2, 9, 9, 4, 8, 3, 2, 1), (3, 7, 6, 8, 4, 1, 1, 0),
                                            (2, 8, 8, 6, 10, 1, 2, 0), (2, 17, 15, 3, 4, 7, 2, 3)]:
        x = rnd(n, c, h, w).requires_grad_()
        wt = rnd(k, c, r, r).requires_grad_()
        y = Fn.conv2d(x, wt, stride=stride, padding=pad)
        up = rnd(*y.shape)
        y.backward(up)
        dy = up.permute(0, 2, 3, 1).contiguous().to(dev)
        xk = x.detach().permute(0, 2, 3, 1).contiguous().to(dev)
        wk = wt.detach().permute(0, 2, 3, 1).contiguous().to(dev)
        dx = torch.full((n, h, w, c), 0.5, device=dev)
        C.vlnce_conv2d_ref_dgrad(ptr(dy), ptr(wk), ptr(dx), n, h, w, c, k, r, r, stride, pad, 1, st)
        torch.testing.assert_close(dx.cpu() - 0.5, x.grad.permute(0, 2, 3, 1), rtol=1e-4, atol=1e-4)
        dw = torch.empty(k, r, r, c, device=dev)
        C.vlnce_conv2d_ref_wgrad(ptr(xk), lib.code(xk), ptr(dy), ptr(dw), n, h, w, c, k, r, r, stride, pad, 0.5, 0, st)
        torch.testing.assert_close(dw.cpu(), 0.5 * wt.grad.permute(0, 2, 3, 1), rtol=1e-4, atol=1e-4)
    # u8 image operand of the RGB stem
    img = torch.randint(0, 256, (2, 12, 12, 3), generator=g0, dtype=torch.uint8)
    wt = rnd(4, 3, 7, 7).requires_grad_()
    y = Fn.conv2d(img.float().permute(0, 3, 1, 2), wt, stride=2, padding=3)
    up = rnd(*y.shape)
    y.backward(up)
    dw = torch.empty(4, 7, 7, 3, device=dev)
    imgd, dy = img.to(dev), up.permute(0, 2, 3, 1).contiguous().to(dev)
    C.vlnce_conv2d_ref_wgrad(ptr(imgd), lib.code(imgd), ptr(dy), ptr(dw), 2, 12, 12, 3, 4, 7, 7, 2, 3, 1.0, 0, st)
    torch.testing.assert_close(dw.cpu(), wt.grad.permute(0, 2, 3, 1), rtol=1e-4, atol=2e-2)
    # ---- GroupNorm backward (channels-last raw conv output)
    for (n, hw, c, groups) in [(3, 20, 32, 16), (2, 9, 128, 1), (2, 16, 24, 4)]:
        x = (rnd(n, c, hw) * 2 + 0.3).requires_grad_()
        gn = torch.nn.GroupNorm(groups, c)
        gn.weight.data.copy_(rnd(c))
        gn.bias.data.copy_(rnd(c))
        y = gn(x)
        up = rnd(*y.shape)
        y.backward(up)
        xk, dy = x.detach().permute(0, 2, 1).contiguous().to(dev), up.permute(0, 2, 1).contiguous().to(dev)
        gam = gn.weight.detach().to(dev)
        dx, part = torch.empty(n, hw, c, device=dev), torch.empty(n, 2 * c, device=dev)
        C.vlnce_groupnorm_bwd(ptr(xk), ptr(dy), ptr(gam), ptr(dx), ptr(part), n, hw, c, groups, float(gn.eps), st)
        torch.testing.assert_close(dx.cpu(), x.grad.permute(0, 2, 1), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(part.cpu().sum(0)[:c], gn.weight.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(part.cpu().sum(0)[c:], gn.bias.grad, rtol=1e-4, atol=1e-4)
    # ---- BatchNorm backward, batch statistics and running statistics
    for training in (1, 0):
        rows, c = 37, 12
        x = (rnd(rows, c) * 1.5 - 0.2).requires_grad_()
        bn = torch.nn.BatchNorm1d(c)
        bn.weight.data.copy_(rnd(c))
        bn.running_mean.copy_(rnd(c) * 0.1)
        bn.running_var.copy_(torch.rand(c, generator=g0) + 0.5)
        rm, rv = bn.running_mean.clone(), bn.running_var.clone()
        bn.train(bool(training))
        y = bn(x)
        up = rnd(*y.shape)
        y.backward(up)
        t = lambda v: v.detach().contiguous().to(dev)  # noqa: E731
        dx, dg, db = torch.empty(rows, c, device=dev), torch.empty(c, device=dev), torch.empty(c, device=dev)
        xk, dyk, gk, rmk, rvk = t(x), t(up), t(bn.weight), t(rm), t(rv)
        C.vlnce_batchnorm_bwd(ptr(xk), ptr(dyk), ptr(gk), ptr(rmk), ptr(rvk), ptr(dx), ptr(dg), ptr(db), rows, c,
                              float(bn.eps), training, st)
        torch.testing.assert_close(dx.cpu(), x.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(dg.cpu(), bn.weight.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(db.cpu(), bn.bias.grad, rtol=1e-4, atol=1e-5)
    # ---- max pool (with exact ties: ReLU outputs) and adaptive average pools (overlapping bins: 7 -> 4)
    x = rnd(2, 5, 9, 8).clamp_min(0).requires_grad_()
    y = Fn.max_pool2d(x, 3, 2, 1)
    up = rnd(*y.shape)
    y.backward(up)
    xk, dy = x.detach().permute(0, 2, 3, 1).contiguous().to(dev), up.permute(0, 2, 3, 1).contiguous().to(dev)
    dx = torch.full((2, 9, 8, 5), 7.0, device=dev)
    C.vlnce_maxpool3x3s2_bwd(ptr(xk), ptr(dy), ptr(dx), 2, 9, 8, 5, st)
    torch.testing.assert_close(dx.cpu(), x.grad.permute(0, 2, 3, 1), rtol=1e-5, atol=1e-6)
    for (h, w, o) in [(7, 7, 4), (8, 8, 4), (5, 6, 1)]:
        x = rnd(2, 6, h, w).requires_grad_()
        y = Fn.adaptive_avg_pool2d(x, (o, o))
        up = rnd(*y.shape)
        y.backward(up)
        ld = 9
        dout = torch.zeros(2, o * o, ld)
        dout[:, :, :6] = up.flatten(2).permute(0, 2, 1)
        dout = dout.to(dev)
        dx = torch.empty(2, h, w, 6, device=dev)
        C.vlnce_adaptive_avgpool_bwd(ptr(dout), ld, ptr(dx), 2, h, w, 6, o, st)
        torch.testing.assert_close(dx.cpu(), x.grad.permute(0, 2, 3, 1), rtol=1e-5, atol=1e-6)


def test_trunk_backward_entry_points_host_model():
    from abi_emulator import Emulator

    with lib.dry_run(executor=Emulator()):
        run("cpu", lib.C)
