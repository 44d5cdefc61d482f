"""GPU bring-up of the individual kernels against torch references (run on the B200 box via gpurun).

    python scripts/bringup.py list            -> case names
    python scripts/bringup.py <case> [...]    -> run cases in this process
    python scripts/bringup.py all             -> run every case, each in its own subprocess with a timeout

Diagnostics are verbose on purpose: this is the tool for finding descriptor / layout mistakes with few GPU calls.
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def rel_err(a, b):
    a = a.double().cpu()
    b = b.double().cpu()
    denom = b.abs().max().item()
    return ((a - b).abs().max().item() / (denom if denom > 0 else 1.0)), denom


def report(name, got, ref, tol):
    e, d = rel_err(got, ref)
    ok = e <= tol and torch.isfinite(got.float()).all().item()
    print("  %-44s max|d|/max|ref| = %.3e (ref max %.3e) tol %.1e  %s" % (name, e, d, tol, "OK" if ok else "FAIL"))
    if not ok:
        g = got.double().cpu()
        r = ref.double().cpu()
        bad = ((g - r).abs() > tol * max(d, 1e-30))
        print("    mismatching elements: %d / %d" % (bad.sum().item(), bad.numel()))
        idx = bad.nonzero()[:8]
        for i in idx:
            t = tuple(i.tolist())
            print("     at %s got %.6g ref %.6g" % (t, g[t].item(), r[t].item()))
        if g.dim() == 4 and g.shape[-1] % 8 == 0:
            print("    bad count per channel-block of 8 (first 16):", bad.sum(dim=(0, 1, 2)).view(-1, 8).sum(1)[:16].tolist())
            print("    bad count per row y (first 16):", bad.sum(dim=(0, 2, 3))[:16].tolist())
            print("    bad count per col x (first 16):", bad.sum(dim=(0, 1, 3))[:16].tolist())
        if g.dim() == 2 and g.shape[0] % 8 == 0 and g.shape[1] % 8 == 0:
            print("    bad rows (first 16 blocks of 8):", bad.sum(1).view(-1)[:128].view(-1, 8).sum(1).tolist())
            print("    bad cols (first 16 blocks of 8):", bad.sum(0).view(-1)[:128].view(-1, 8).sum(1).tolist())
    return ok


def ref_conv(x, w_hwio, bias=None, relu=False):
    """NHWC x, HWIO w -> NHWC, stride 1 SAME, computed in fp64 on the GPU."""
    k = w_hwio.shape[0]
    y = F.conv2d(x.double().permute(0, 3, 1, 2), w_hwio.double().permute(3, 2, 0, 1), padding=k // 2)
    y = y.permute(0, 2, 3, 1)
    if bias is not None:
        y = y + bias.double()
    if relu:
        y = y.clamp_min(0)
    return y


def conv_case(N, H, W, cin, cout, k, dtype, tol, seed=0, force_bn=0, force_splits=0, flags_extra=0, nseg=1):
    from fcn8s_tensorflow_b200 import ops
    torch.manual_seed(seed)
    dev = torch.device("cuda")
    tdt = ops.torch_dtype(dtype)
    x = torch.randn(N, H, W, cin, device=dev).to(tdt)
    w = (torch.randn(k, k, cin, cout, device=dev) / (k * k * cin) ** 0.5)
    b = torch.randn(cout, device=dev)
    if nseg == 3:
        wp, wlo = ops.pack_weights(w, k, cin, cout, 0, dtype, split=True)
        xh, xl = ops.split_tf32(x)
        y = ops.conv_gemm(xh, wp, cout, k, bias=b, flags=ops.EPI_BIAS | ops.EPI_RELU, x_lo=xl, wp_lo=wlo,
                          force_bn=force_bn, force_splits=force_splits)
        wq = w
    else:
        wp, _ = ops.pack_weights(w, k, cin, cout, 0, dtype)
        y = ops.conv_gemm(x, wp, cout, k, bias=b, flags=ops.EPI_BIAS | ops.EPI_RELU, force_bn=force_bn,
                          force_splits=force_splits)
        wq = w.to(tdt) if dtype == ops.BF16 else w
    torch.cuda.synchronize()
    ref = ref_conv(x, wq, b, relu=True)
    return report("conv N%d %dx%d Cin%d Cout%d k%d dt%d bn%d sp%d seg%d" %
                  (N, H, W, cin, cout, k, dtype, force_bn, force_splits, nseg), y, ref, tol)


@case
def conv_bf16_1x1_single_tile():
    return conv_case(1, 8, 16, 64, 64, 1, 0, 1e-2)


@case
def conv_bf16_1x1_k128():
    return conv_case(1, 8, 16, 128, 64, 1, 0, 1e-2)


@case
def conv_bf16_3x3_single_tile():
    return conv_case(1, 8, 16, 64, 64, 3, 0, 1e-2)


@case
def conv_bf16_3x3_bn128_bn256():
    ok = conv_case(2, 16, 32, 128, 128, 3, 0, 1e-2)
    ok &= conv_case(2, 16, 32, 128, 256, 3, 0, 1e-2)
    ok &= conv_case(2, 16, 32, 64, 512, 3, 0, 1e-2, force_bn=256)
    return ok


@case
def conv_bf16_persistent_many_tiles():
    ok = conv_case(2, 64, 128, 64, 64, 3, 0, 1e-2)      # 128 m-tiles
    ok &= conv_case(4, 64, 128, 64, 128, 3, 0, 1e-2)    # 256 m-tiles > 148 CTAs: persistent loop + TMEM double buffer
    ok &= conv_case(3, 40, 72, 128, 256, 3, 0, 1e-2)    # ragged: partial tiles in every dim
    return ok


@case
def conv_bf16_splitk_and_7x7():
    ok = conv_case(1, 8, 16, 512, 256, 3, 0, 1e-2, force_splits=3)
    ok &= conv_case(2, 4, 8, 512, 256, 7, 0, 1e-2)       # fc6 geometry (heuristic split-K)
    ok &= conv_case(1, 1, 2, 512, 512, 7, 0, 1e-2)       # tiny spatial: box larger than the tensor
    return ok


@case
def conv_tf32_basic():
    ok = conv_case(1, 8, 16, 32, 64, 1, 1, 2e-3)
    ok &= conv_case(1, 8, 16, 64, 64, 3, 1, 2e-3)
    ok &= conv_case(2, 16, 32, 128, 256, 3, 1, 2e-3)
    return ok


@case
def conv_tf32x3():
    ok = conv_case(1, 8, 16, 64, 64, 3, 1, 2e-6, nseg=3)
    ok &= conv_case(2, 16, 32, 128, 256, 3, 1, 2e-6, nseg=3)
    return ok


@case
def tf32_truncation_probe():
    """Does kind::tf32 truncate or round the low 13 mantissa bits of its fp32 operands?  (Decides whether the hi part
    of the 3xTF32 split has to be materialised.)"""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(11)
    x = torch.randn(1, 8, 16, 64, device=dev)
    w = torch.randn(3, 3, 64, 64, device=dev) / 24.0
    wp, _ = ops.pack_weights(w, 3, 64, 64, 0, 1)
    xh, _ = ops.split_tf32(x)
    wph = (wp.view(torch.int32) & -8192).view(torch.float32)
    y_full = ops.conv_gemm(x, wp, 64, 3)
    y_trunc = ops.conv_gemm(xh, wph.contiguous(), 64, 3)
    torch.cuda.synchronize()
    same = torch.equal(y_full, y_trunc)
    print("  tf32 operands truncated by hardware (bitwise equal outputs): %s; max diff %.3e" %
          (same, (y_full - y_trunc).abs().max().item()))
    return True


@case
def conv_epilogues():
    """dgrad-style epilogue: residual add + relu mask with scale; and dropout."""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    ok = True
    for dtype, tol in ((0, 1e-2), (1, 2e-3)):
        torch.manual_seed(1)
        tdt = ops.torch_dtype(dtype)
        N, H, W, cin, cout, k = 2, 16, 16, 128, 128, 3
        x = torch.randn(N, H, W, cin, device=dev).to(tdt)
        w = torch.randn(k, k, cin, cout, device=dev) / (k * k * cin) ** 0.5
        msrc = torch.randn(N, H, W, cout, device=dev).to(tdt)
        res = torch.randn(N, H, W, cout, device=dev).to(tdt)
        wp, _ = ops.pack_weights(w, k, cin, cout, 0, dtype)
        wq = w.to(tdt) if dtype == 0 else w
        for splits in (0, 2):
            y = ops.conv_gemm(x, wp, cout, k, flags=ops.EPI_MASK | ops.EPI_RESIDUAL, mask_src=msrc, residual=res,
                              mask_scale=2.0, force_splits=splits)
            ref = (ref_conv(x, wq) + res.double()) * (msrc.double() > 0) * 2.0
            ok &= report("epilogue mask+residual dt%d splits%d" % (dtype, splits), y, ref, tol)
            b = torch.randn(cout, device=dev)
            y = ops.conv_gemm(x, wp, cout, k, bias=b, flags=ops.EPI_BIAS | ops.EPI_RELU | ops.EPI_DROPOUT,
                              keep_prob=0.5, seed=1234, force_splits=splits)
            from fcn8s_tensorflow_b200.rng import dropout_keep_mask
            keep = dropout_keep_mask(1234, y.numel(), 0.5).view(y.shape).to(dev)
            ref = ref_conv(x, wq, b, relu=True) * keep.double() * 2.0
            ok &= report("epilogue dropout dt%d splits%d (keep frac %.3f)" % (dtype, splits, keep.float().mean().item()),
                         y, ref, tol)
    return ok


def dgrad_case(N, H, W, cin, cout, k, dtype, tol):
    from fcn8s_tensorflow_b200 import ops
    torch.manual_seed(3)
    dev = torch.device("cuda")
    tdt = ops.torch_dtype(dtype)
    w = torch.randn(k, k, cin, cout, device=dev) / (k * k * cout) ** 0.5
    dy = torch.randn(N, H, W, cout, device=dev).to(tdt)
    wp, _ = ops.pack_weights(w, k, cin, cout, 1, dtype)
    dx = ops.conv_gemm(dy, wp, cin, k)
    wq = (w.to(tdt) if dtype == 0 else w).double()
    ref = F.conv_transpose2d(dy.double().permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
    return report("dgrad N%d %dx%d Cin%d Cout%d k%d dt%d" % (N, H, W, cin, cout, k, dtype), dx, ref, tol)


@case
def dgrad_pack_and_conv():
    ok = dgrad_case(2, 16, 32, 64, 128, 3, 0, 1e-2)
    ok &= dgrad_case(1, 4, 8, 512, 256, 7, 0, 1e-2)
    ok &= dgrad_case(2, 16, 32, 64, 128, 3, 1, 2e-3)
    return ok


def wgrad_case(N, H, W, cin, cout, k, dtype, tol, force_splits=0, force_bn=0, rows_valid=0, nseg=1):
    from fcn8s_tensorflow_b200 import ops
    torch.manual_seed(4)
    dev = torch.device("cuda")
    tdt = ops.torch_dtype(dtype)
    x = torch.randn(N, H, W, cin, device=dev).to(tdt)
    dy = torch.randn(N, H, W, cout, device=dev).to(tdt)
    rows = k * k * cin
    rv = rows_valid if rows_valid else rows
    dw = torch.full((rv, cout), float("nan"), device=dev)
    if nseg == 3:
        xh, xl = ops.split_tf32(x)
        dh, dl = ops.split_tf32(dy)
        ops.wgrad_gemm(xh, dh, k, dw, rows_valid=rows_valid, x_lo=xl, dy_lo=dl, force_splits=force_splits,
                       force_bn=force_bn)
    else:
        ops.wgrad_gemm(x, dy, k, dw, rows_valid=rows_valid, force_splits=force_splits, force_bn=force_bn)
    torch.cuda.synchronize()
    # reference: dW[kh,kw,ci,co] = sum x[n,y+kh-p,x+kw-p,ci] dy[n,y,x,co]
    xp = F.pad(x.double().permute(0, 3, 1, 2), (k // 2,) * 4)
    cols = F.unfold(xp, k).view(N, cin, k * k, H * W)             # [N, ci, tap, P]
    ref = torch.einsum("nctp,npo->tco", cols, dy.double().reshape(N, H * W, cout)).reshape(rows, cout)[:rv]
    return report("wgrad N%d %dx%d Cin%d Cout%d k%d dt%d sp%d bn%d seg%d" %
                  (N, H, W, cin, cout, k, dtype, force_splits, force_bn, nseg), dw, ref, tol)


@case
def wgrad_bf16_1x1_single():
    return wgrad_case(1, 8, 8, 128, 64, 1, 0, 1e-2)


@case
def wgrad_bf16_1x1_m64():
    return wgrad_case(1, 8, 8, 64, 64, 1, 0, 1e-2)


@case
def wgrad_bf16_3x3():
    ok = wgrad_case(1, 8, 8, 64, 64, 3, 0, 1e-2)
    ok &= wgrad_case(2, 16, 32, 128, 256, 3, 0, 1e-2)
    ok &= wgrad_case(3, 20, 36, 64, 128, 3, 0, 1e-2)
    return ok


@case
def wgrad_bf16_splits_7x7_rows():
    ok = wgrad_case(2, 32, 32, 64, 64, 3, 0, 1e-2, force_splits=1)
    ok &= wgrad_case(2, 32, 32, 64, 64, 3, 0, 1e-2, force_splits=5)
    ok &= wgrad_case(2, 4, 8, 512, 256, 7, 0, 1e-2)
    ok &= wgrad_case(2, 16, 16, 64, 64, 1, 0, 1e-2, rows_valid=27)
    return ok


@case
def wgrad_tf32():
    ok = wgrad_case(1, 8, 8, 32, 64, 1, 1, 2e-3)
    ok &= wgrad_case(2, 16, 32, 128, 256, 3, 1, 2e-3)
    ok &= wgrad_case(2, 16, 32, 64, 128, 3, 1, 2e-6, nseg=3)
    return ok


@case
def elementwise_kernels():
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    ok = True
    torch.manual_seed(5)
    # preprocess
    img = torch.randint(0, 256, (2, 10, 14, 3), dtype=torch.uint8, device=dev)
    mean = torch.tensor([103.939, 116.779, 123.68], device=dev, dtype=torch.float64)
    bgr = img.double().flip(-1) - mean
    xp = F.pad(bgr.permute(0, 3, 1, 2), (1, 1, 1, 1))
    cols = F.unfold(xp, 3).view(2, 3, 9, 10, 14).permute(0, 3, 4, 2, 1).reshape(2, 10, 14, 27)
    for dtype, tol in ((0, 4e-3), (1, 1e-7)):
        out = ops.preprocess_im2col(img, dtype)
        ok &= report("preprocess dt%d" % dtype, out[..., :27], cols, tol)
        ok &= bool((out[..., 27:] == 0).all().item())
    # pool fwd / bwd (odd sizes)
    for dtype in (0, 1):
        tdt = ops.torch_dtype(dtype)
        x = torch.randn(2, 9, 13, 64, device=dev).clamp_min(0).to(tdt)
        y = ops.maxpool_fwd(x)
        xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
        yr = F.max_pool2d(xr, 2, 2, ceil_mode=True)
        ok &= report("maxpool fwd dt%d" % dtype, y, yr.permute(0, 2, 3, 1), 0.0)
        dy = torch.randn_like(y.float()).to(tdt)
        dx = ops.maxpool_bwd(x, dy)
        yr.backward(dy.double().permute(0, 3, 1, 2))
        ref = xr.grad.permute(0, 2, 3, 1) * (x.double() > 0)
        ok &= report("maxpool bwd dt%d" % dtype, dx, ref, 0.0)
        if dtype == 0:
            # bias gradient of the producer, accumulated by the same kernel, for every VGG width (the lanes of a warp
            # that share a channel vector are pre-reduced with shuffles when C < 256)
            for Cc in (64, 128, 256, 512):
                xc = torch.randn(2, 9, 13, Cc, device=dev).clamp_min(0).to(tdt)
                dyc = torch.randn(2, 5, 7, Cc, device=dev).to(tdt)
                db = torch.full((Cc,), 0.5, device=dev)
                dxc = ops.maxpool_bwd(xc, dyc, db=db)
                ok &= report("maxpool bwd db C%d" % Cc, db, 0.5 + dxc.double().sum((0, 1, 2)), 1e-5)
            try:     # 192 / 8 = 24 channel vectors do not divide the 256-thread grid stride: refused, not mis-summed
                ops.maxpool_bwd(torch.zeros(1, 4, 4, 192, device=dev).to(tdt), torch.zeros(1, 2, 2, 192, device=dev).to(tdt),
                                db=torch.zeros(192, device=dev))
                ok = False
                print("  maxpool bwd db C192 was not refused: FAIL")
            except Exception:  # noqa: BLE001
                pass
        # bias grad
        for Cc in (64, 512, 4096):
            g = torch.randn(1000, Cc, device=dev).to(tdt)
            db = torch.empty(Cc, device=dev)
            ops.bias_grad(g, db)
            ok &= report("bias_grad C%d dt%d" % (Cc, dtype), db, g.double().sum(0), 1e-5)
    # adam
    n = 1000003
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    m = torch.randn(n, device=dev) * 0.1
    v = torch.rand(n, device=dev) * 0.1
    pr, mr, vr = p.double(), m.double(), v.double()
    gs = 0.5
    mr = 0.9 * mr + 0.1 * g.double() * gs
    vr = 0.999 * vr + 0.001 * (g.double() * gs) ** 2
    pr = pr - 1e-3 * mr / (vr.sqrt() + 1e-8)
    ops.adam(p, g, m, v, 1e-3, grad_scale=gs)
    ok &= report("adam p", p, pr, 1e-6)
    ok &= report("adam m", m, mr, 1e-6)
    ok &= report("adam v", v, vr, 1e-6)
    # l2
    w = torch.randn(5000, device=dev)
    g2 = torch.zeros(5000, device=dev)
    ls = torch.zeros(1, device=dev)
    ops.l2_reg(w, g2, ls, 0.01)
    ok &= report("l2 grad", g2, 0.01 * w.double(), 1e-6)
    ok &= report("l2 loss", ls, (0.005 * (w.double() ** 2).sum()).view(1), 1e-5)
    return ok


# ---------------------------------------------------------------------------------------------------------------------
# bf16 / bf16 hi-lo pair modes that read the TF-layout (HWIO) weight shadow in place (no packing)
def _shadow(w, pair):
    from fcn8s_tensorflow_b200 import ops
    hi = torch.empty(w.numel(), dtype=torch.bfloat16, device=w.device)
    lo = torch.empty_like(hi) if pair else None
    ops.shadow_weights(w.reshape(-1), hi, lo)
    return hi.view(w.shape), (lo.view(w.shape) if pair else None)


def hwio_conv_case(N, H, W, cin, cout, k, pair, tol, force_bn=0, force_splits=0, algo=0):
    """fprop (w_mode 1) with bias+ReLU, and dgrad (w_mode 2), against fp64 references on the operands' exact values."""
    from fcn8s_tensorflow_b200 import ops
    torch.manual_seed(21)
    dev = torch.device("cuda")
    w = torch.randn(k, k, cin, cout, device=dev) / (k * k * cin) ** 0.5
    b = torch.randn(cout, device=dev)
    wh, wl = _shadow(w, pair)
    wq = (wh.double() + wl.double()) if pair else wh.double()
    x32 = torch.randn(N, H, W, cin, device=dev)
    x = ops.to_pair(x32) if pair else x32.to(torch.bfloat16)
    xq = ops.from_pair(x).double() if pair else x.double()
    y = ops.conv_gemm(x, wh, cout, k, bias=b, flags=ops.EPI_BIAS | ops.EPI_RELU, wp_lo=wl, pair=pair, w_mode=1,
                      force_bn=force_bn, force_splits=force_splits, algo=algo)
    torch.cuda.synchronize()
    ref = ref_conv(xq, wq, b, relu=True)
    got = ops.from_pair(y) if pair else y
    tag = "N%d %dx%d Cin%d Cout%d k%d pair%d bn%d sp%d algo%d" % (N, H, W, cin, cout, k, pair, force_bn, force_splits,
                                                                   algo)
    ok = report("hwio fprop " + tag, got, ref, tol)
    # dgrad: dy has cout channels, result cin channels
    dy32 = torch.randn(N, H, W, cout, device=dev)
    dy = ops.to_pair(dy32) if pair else dy32.to(torch.bfloat16)
    dyq = ops.from_pair(dy).double() if pair else dy.double()
    dx = ops.conv_gemm(dy, wh, cin, k, wp_lo=wl, pair=pair, w_mode=2, force_bn=force_bn if cin % max(force_bn, 1) == 0 else 0,
                       force_splits=force_splits, algo=algo)
    torch.cuda.synchronize()
    refd = F.conv_transpose2d(dyq.permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), padding=k // 2).permute(0, 2, 3, 1)
    ok &= report("hwio dgrad " + tag, ops.from_pair(dx) if pair else dx, refd, tol)
    return ok


@case
def hwio_bf16():
    ok = hwio_conv_case(1, 8, 16, 64, 64, 1, False, 1e-2)
    ok &= hwio_conv_case(1, 8, 16, 64, 64, 3, False, 1e-2)
    ok &= hwio_conv_case(2, 16, 32, 128, 256, 3, False, 1e-2)
    ok &= hwio_conv_case(2, 16, 32, 64, 512, 3, False, 1e-2, force_bn=256)
    ok &= hwio_conv_case(2, 16, 32, 256, 128, 3, False, 1e-2, force_bn=128)
    ok &= hwio_conv_case(1, 4, 8, 512, 256, 7, False, 1e-2)              # split-K path
    ok &= hwio_conv_case(3, 20, 36, 64, 128, 3, False, 1e-2)             # ragged tiles
    return ok


@case
def hwio_pair():
    # hi/lo pair operands: products carry ~2^-17 relative error per element, reductions average it down
    ok = hwio_conv_case(1, 8, 16, 64, 64, 1, True, 2e-5)
    ok &= hwio_conv_case(2, 16, 32, 128, 256, 3, True, 2e-5)
    ok &= hwio_conv_case(2, 16, 32, 64, 512, 3, True, 2e-5, force_bn=256)
    ok &= hwio_conv_case(1, 4, 8, 512, 256, 7, True, 2e-5)               # split-K path with pair epilogue
    ok &= hwio_conv_case(3, 20, 36, 64, 128, 3, True, 2e-5)
    return ok


@case
def cta_pair_kernels():
    """CTA-pair (cta_group::2) variants of the per-tap conv kernel and of the filter-gradient kernel: odd numbers of M
    tiles (the last pair has a partner that computes on zero-filled rows), ragged tiles, split-K, bf16 and hi/lo pair
    operands -- against the fp64 references, and bit-for-bit against the single-CTA kernels (same accumulation order)."""
    from fcn8s_tensorflow_b200 import _capi, ops
    lib = _capi.load()
    ok = hwio_conv_case(3, 20, 36, 256, 256, 3, False, 1e-2, algo=1)      # 17 M tiles, ragged, Cin = Cout = 256
    ok &= hwio_conv_case(1, 8, 16, 256, 512, 3, False, 1e-2, algo=1)      # a single M tile, two N tiles
    ok &= hwio_conv_case(3, 20, 36, 128, 256, 3, True, 2e-5, algo=1)      # hi/lo pair operands (3 segments)
    ok &= hwio_conv_case(1, 4, 8, 512, 256, 7, False, 1e-2)               # split-K partials
    ok &= wgrad_case(3, 20, 36, 128, 256, 3, 0, 1e-2)                     # 9 M tiles (odd), 256 columns
    ok &= wgrad_case(2, 16, 32, 256, 512, 3, 0, 1e-2)                     # 18 M tiles, two N tiles
    ok &= wgrad_case(1, 8, 16, 64, 256, 1, 0, 1e-2)                       # half an M tile
    # pair vs single-CTA kernel on the same inputs
    torch.manual_seed(33)
    dev = torch.device("cuda")
    w = torch.randn(3, 3, 256, 256, device=dev) / 48.0
    wh, _ = _shadow(w, False)
    x = torch.randn(3, 20, 36, 256, device=dev).to(torch.bfloat16)
    y_pair = ops.conv_gemm(x, wh, 256, 3, w_mode=1, algo=1)
    lib.fcn8_debug_set(5, 1)
    try:
        y_single = ops.conv_gemm(x, wh, 256, 3, w_mode=1, algo=1)
    finally:
        lib.fcn8_debug_set(5, 0)
    torch.cuda.synchronize()
    same = bool(torch.equal(y_pair, y_single))
    print("  %-58s %s" % ("pair kernel == single-CTA kernel, bit for bit", "OK" if same else "FAIL"))
    return ok and same


@case
def halo_conv():
    """Halo-tile kernel (algo 2): shifted UMMA descriptors into one activation patch, against the same references."""
    ok = True
    for algo in (2,):   # (a variant that set the descriptors' base-offset field to (start >> 7) & 7 gave wrong results:
        #                  the 128B swizzle is a function of the absolute shared-memory address)
        print(" -- algo %d" % algo)
        ok_a = hwio_conv_case(1, 16, 8, 64, 64, 3, False, 1e-2, algo=algo)          # exactly one tile
        ok_a &= hwio_conv_case(2, 32, 32, 64, 64, 3, False, 1e-2, algo=algo)
        ok_a &= hwio_conv_case(2, 32, 32, 128, 128, 3, False, 1e-2, algo=algo)       # 2 channel blocks
        ok_a &= hwio_conv_case(3, 20, 36, 64, 128, 3, False, 1e-2, algo=algo)        # ragged tiles
        ok_a &= hwio_conv_case(1, 48, 40, 128, 256, 3, False, 1e-2, algo=algo)
        ok_a &= hwio_conv_case(2, 20, 28, 64, 128, 3, True, 2e-5, algo=algo)         # hi/lo pairs
        ok_a &= hwio_conv_case(2, 36, 20, 64, 64, 3, True, 2e-5, algo=algo)          # pairs, weights resident in smem
        ok_a &= hwio_conv_case(1, 40, 24, 128, 64, 3, False, 1e-2, algo=algo)        # resident, 2 channel blocks
        print(" -- algo %d -> %s" % (algo, "PASS" if ok_a else "FAIL"))
        if algo == 2:
            ok = ok_a     # algo 3 (base-offset variant) is informational
    return ok


@case
def pair_epilogues_and_wgrad():
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(22)
    ok = True
    N, H, W, cin, cout, k = 2, 16, 32, 128, 128, 3
    w = torch.randn(k, k, cin, cout, device=dev) / (k * k * cin) ** 0.5
    wh, wl = _shadow(w, True)
    wq = wh.double() + wl.double()
    x = ops.to_pair(torch.randn(N, H, W, cin, device=dev))
    xq = ops.from_pair(x).double()
    msrc = ops.to_pair(torch.randn(N, H, W, cout, device=dev).clamp_min(0))
    res = ops.to_pair(torch.randn(N, H, W, cout, device=dev))
    base = ref_conv(xq, wq)
    y = ops.conv_gemm(x, wh, cout, k, flags=ops.EPI_MASK, mask_src=msrc, mask_scale=2.0, wp_lo=wl, pair=True, w_mode=1)
    ok &= report("pair mask epilogue", ops.from_pair(y), base * (ops.from_pair(msrc).double() > 0) * 2.0, 2e-5)
    y = ops.conv_gemm(x, wh, cout, k, flags=ops.EPI_RESIDUAL, residual=res, wp_lo=wl, pair=True, w_mode=1)
    ok &= report("pair residual epilogue", ops.from_pair(y), base + ops.from_pair(res).double(), 2e-5)
    # wgrad on pairs
    for (n_, h_, w_, ci, co, kk, rv) in ((2, 16, 32, 128, 256, 3, 0), (3, 20, 36, 64, 64, 3, 0), (2, 16, 16, 64, 64, 1, 27)):
        xx = ops.to_pair(torch.randn(n_, h_, w_, ci, device=dev))
        dy = ops.to_pair(torch.randn(n_, h_, w_, co, device=dev))
        rows = kk * kk * ci
        rvv = rv if rv else rows
        dw = torch.full((rvv, co), float("nan"), device=dev)
        ops.wgrad_gemm(xx, dy, kk, dw, rows_valid=rv, pair=True)
        xpd = F.pad(ops.from_pair(xx).double().permute(0, 3, 1, 2), (kk // 2,) * 4)
        cols = F.unfold(xpd, kk).view(n_, ci, kk * kk, h_ * w_)
        ref = torch.einsum("nctp,npo->tco", cols, ops.from_pair(dy).double().reshape(n_, h_ * w_, co)).reshape(rows, co)[:rvv]
        ok &= report("pair wgrad Cin%d Cout%d k%d" % (ci, co, kk), dw, ref, 2e-5)
    return ok


@case
def pair_elementwise():
    """Feed, pooling, bias-gradient and the Adam shadow in the bf16 hi/lo pair format."""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(23)
    ok = True
    img = torch.randint(0, 256, (2, 10, 14, 3), dtype=torch.uint8, device=dev)
    mean = torch.tensor([103.939, 116.779, 123.68], device=dev, dtype=torch.float64)
    bgr = img.double().flip(-1) - mean
    xp = F.pad(bgr.permute(0, 3, 1, 2), (1, 1, 1, 1))
    cols = F.unfold(xp, 3).view(2, 3, 9, 10, 14).permute(0, 3, 4, 2, 1).reshape(2, 10, 14, 27)
    out = ops.preprocess_im2col(img, ops.BF16X2)
    ok &= report("preprocess pair", ops.from_pair(out)[..., :27], cols, 2e-5)
    ok &= bool((ops.from_pair(out)[..., 27:] == 0).all().item())
    x = ops.to_pair(torch.randn(2, 9, 13, 64, device=dev).clamp_min(0))
    y = ops.maxpool_fwd(x, pair=True)
    xr = ops.from_pair(x).double().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2, ceil_mode=True)
    ok &= report("maxpool fwd pair", ops.from_pair(y), yr.permute(0, 2, 3, 1), 0.0)
    dy = ops.to_pair(torch.randn(2, 5, 7, 64, device=dev))
    dx = ops.maxpool_bwd(x, dy, pair=True)
    yr.backward(ops.from_pair(dy).double().permute(0, 3, 1, 2))
    ok &= report("maxpool bwd pair", ops.from_pair(dx), xr.grad.permute(0, 2, 3, 1) * (ops.from_pair(x).double() > 0), 0.0)
    for Cc in (64, 128, 512):
        xc = ops.to_pair(torch.randn(2, 9, 13, Cc, device=dev).clamp_min(0))
        dyc = ops.to_pair(torch.randn(2, 5, 7, Cc, device=dev))
        db = torch.zeros(Cc, device=dev)
        dxc = ops.maxpool_bwd(xc, dyc, pair=True, db=db)
        ok &= report("maxpool bwd pair db C%d" % Cc, db, ops.from_pair(dxc).double().sum((0, 1, 2)), 1e-5)
    for Cc in (64, 4096):
        g = ops.to_pair(torch.randn(1000, Cc, device=dev))
        db = torch.empty(Cc, device=dev)
        ops.bias_grad(g, db, pair=True)
        ok &= report("bias_grad pair C%d" % Cc, db, ops.from_pair(g).double().sum(0), 1e-5)
    # Adam refreshes the shadow
    n = 100003
    p = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    m = torch.zeros(n, device=dev)
    v = torch.zeros(n, device=dev)
    hi = torch.empty(n, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi)
    ops.adam(p, g, m, v, 1e-3, w_hi=hi, w_lo=lo)
    ok &= bool((hi == p.to(torch.bfloat16)).all().item())
    ok &= bool((lo == (p - hi.float()).to(torch.bfloat16)).all().item())
    ok &= report("shadow hi+lo ~ p", hi.float() + lo.float(), p, 2e-5)
    return ok


# ---------------------------------------------------------------------------------------------------------------------
# Decoder on the tensor cores: score heads, transposed convolutions as phase GEMMs over bf16 hi/lo planes, the loss /
# predictor epilogue, the fused max-pool, the reduced-product modes
def ref_deconv(x, T, bias, s):
    """tf.layers.conv2d_transpose(k = 2s, stride s, 'same') in fp64: x [N,h,w,C], T [2s,2s,Cout,Cin] -> [N,sh,sw,C]."""
    y = F.conv_transpose2d(x.double().permute(0, 3, 1, 2), T.double().permute(3, 2, 0, 1), bias.double(), stride=s,
                           padding=s // 2)
    return y.permute(0, 2, 3, 1)


def _pair_q(t):
    """value an fp32 tensor takes when stored as a bf16 hi/lo pair."""
    hi = t.to(torch.bfloat16)
    return hi.double() + (t - hi.float()).to(torch.bfloat16).double()


@case
def decoder_heads():
    """1x1 score heads through fcn8_conv_gemm / fcn8_wgrad_gemm with the classes padded to 64 columns: forward with
    the skip scale and bias, input gradient (with the fc7 mask), kernel gradient clipped to C columns."""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(41)
    ok = True
    for pair, tol in ((True, 2e-5), (False, 1e-2)):
        for Cc, cin, scale in ((20, 256, 1e-2), (5, 512, 1.0), (2, 4096, 1e-4)):
            N, h, w = 2, 6, 10
            x32 = torch.randn(N, h, w, cin, device=dev).clamp_min(0)
            x = ops.to_pair(x32) if pair else x32.to(torch.bfloat16)
            xq = ops.from_pair(x).double() if pair else x.double()
            K = torch.randn(cin, Cc, device=dev) * 0.05
            b = torch.randn(Cc, device=dev)
            pk = ops.head_pack(K, b, split=pair)
            Kq = (pk["w"].double() + (pk["w_lo"].double() if pair else 0))[:, :Cc]
            nseg = 3 if pair else 1
            S = torch.full((N, h, w, 128), float("nan"), dtype=torch.bfloat16, device=dev)
            ops.conv_gemm(x, pk["w"], 64, 1, bias=pk["bias64"], flags=ops.EPI_BIAS, out=S, wp_lo=pk["w_lo"], pair=pair,
                          out_pair=True, w_mode=1, out_scale=scale, nseg=nseg)
            got = ops.from_pair(S)
            ref = scale * (xq.reshape(-1, cin) @ Kq) + b.double()
            tag = "C%d Cin%d pair%d" % (Cc, cin, pair)
            ok &= report("head fwd " + tag, got[..., :Cc].reshape(-1, Cc), ref, tol)
            ok &= bool((got[..., Cc:] == 0).all().item())
            # backward: ds planes as a strided interior view of a padded tensor (what the engine passes)
            ds32 = torch.randn(N, h, w, Cc, device=dev)
            pad = ops.planes_alloc(N, h + 2, w + 2, dev, 64)
            dsv = tuple(t[:, 1:-1, 1:-1, :] for t in pad)
            dsv[0][..., :Cc] = ds32.to(torch.bfloat16)
            dsv[1][..., :Cc] = (ds32 - dsv[0][..., :Cc].float()).to(torch.bfloat16)
            dsq = (dsv[0].double() + (dsv[1].double() if pair else 0))[..., :Cc]
            dK = torch.full((cin, Cc), float("nan"), device=dev)
            ops.wgrad_gemm(x, dsv[0], 1, dK, dy_lo=dsv[1], pair=pair, dy_pair=False, nseg=nseg, out_cols=Cc,
                           out_scale=scale)
            ok &= report("head dK " + tag, dK, scale * xq.reshape(-1, cin).t() @ dsq.reshape(-1, Cc), tol)
            dx = torch.empty_like(x)
            ops.conv_gemm(dsv[0], pk["w"], cin, 1, x_lo=dsv[1] if pair else None, wp_lo=pk["w_lo"], w_mode=2, out=dx,
                          out_pair=pair, out_scale=scale, nseg=nseg, cin=64, flags=ops.EPI_MASK, mask_src=x,
                          mask_scale=2.0)
            refdx = scale * (dsq.reshape(-1, Cc) @ Kq.t()).reshape(xq.shape) * (xq > 0) * 2.0
            ok &= report("head dx " + tag, ops.from_pair(dx) if pair else dx, refdx, tol)
    return ok


def deconv_case(Cc, s, N, h, w, nseg, tol):
    """fwd (s = 2: dense + skip; s = 8: every output of the loss epilogue), dx (+ column sums) and dT of one stage."""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(50 + s + Cc)
    ok = True
    tag = "C%d s%d N%d %dx%d seg%d" % (Cc, s, N, h, w, nseg)
    T = torch.randn(2 * s, 2 * s, Cc, Cc, device=dev) * 0.2
    bias = torch.randn(Cc, device=dev)
    pk = ops.deconv_pack(T, bias, s, split=nseg > 1)
    Tq = _pair_q(T) if nseg == 3 else T.to(torch.bfloat16).double()
    x32 = torch.randn(N, h, w, Cc, device=dev)
    xp = ops.planes_from_float(x32)
    xq = ops.planes_to_float(xp, Cc).double() if nseg == 3 else xp[0][..., :Cc].double()
    H, W = s * h, s * w
    ref = ref_deconv(xq, Tq, bias, s)
    if s == 2:
        skip32 = torch.randn(N, H, W, Cc, device=dev)
        skp = ops.halves(ops.to_pair(F.pad(skip32, (0, 64 - Cc))), 64)    # same geometry as the output planes
        out_t = torch.full((N, H, W, 128), float("nan"), dtype=torch.bfloat16, device=dev)
        out = ops.halves(out_t, 64)
        ops.deconv_fwd(xp, pk, Cc, s, out, skip=skp, nseg=nseg)
        got = ops.planes_to_float(out, 64)
        ok &= report("deconv fwd+skip " + tag, got[..., :Cc], ref + ops.planes_to_float(skp, Cc).double(), tol)
        ok &= bool((got[..., Cc:] == 0).all().item())
    else:
        ids = torch.randint(0, Cc, (N, H, W), device=dev)
        onehot = F.one_hot(ids, Cc).to(torch.uint8)
        # a few rows that are not one-hot: TF's fused op still back-propagates softmax - labels
        onehot[0, 0, :4] = 0
        onehot[0, 1, :4, :min(2, Cc)] = 1
        logits = torch.full((N, H, W, Cc), float("nan"), device=dev)
        sm = torch.full((N, H, W, Cc), float("nan"), device=dev)
        am = torch.full((N, H, W), -1, dtype=torch.int64, device=dev)
        loss = torch.zeros(1, device=dev)
        dbias = torch.zeros(Cc, device=dev)
        conf = torch.zeros((Cc, Cc), dtype=torch.int64, device=dev)
        dz = ops.padded_alloc(N, h, w, 8, dev)
        gs = 1.0 / (N * H * W)
        ops.deconv_loss(xp, pk, Cc, nseg=nseg, labels=onehot, loss_sum=loss, dz_out=dz, dbias=dbias, grad_scale=gs,
                        logits=logits, softmax=sm, argmax=am, conf=conf)
        torch.cuda.synchronize()
        ok &= report("loss-epilogue logits " + tag, logits, ref, tol)
        z = logits.double()    # the rest is checked against the kernel's own logits (tight)
        y = onehot.double()
        lse = torch.logsumexp(z, -1)
        ok &= report("loss sum", loss, ((y.sum(-1) * lse) - (y * z).sum(-1)).sum().view(1), 1e-5)
        ok &= report("softmax", sm, torch.softmax(z, -1), 1e-5)
        ok &= bool(torch.equal(am, z.argmax(-1)))
        dzi = ops.padded_interior(dz, 8)
        dz_ref = (torch.softmax(z, -1) - y) * gs
        ok &= report("dz planes", ops.planes_to_float(dzi, Cc), dz_ref, 2e-5)
        ok &= report("dbias", dbias, dz_ref.sum((0, 1, 2)), 1e-4)
        full = dz[0].float() + dz[1].float()
        border = full.clone()
        border[:, 4:-4, 4:-4, :Cc] = 0
        ok &= bool((border == 0).all().item())
        cm = torch.zeros((Cc, Cc), dtype=torch.int64, device=dev)
        cm.index_put_((onehot.argmax(-1).reshape(-1), am.reshape(-1)), torch.ones(N * H * W, dtype=torch.int64, device=dev),
                      accumulate=True)
        same = bool(torch.equal(cm, conf))
        print("  %-44s %s" % ("confusion matrix (epilogue) exact", "OK" if same else "FAIL"))
        ok &= same
        # each output alone (other pointers NULL)
        am2 = torch.empty_like(am)
        ops.deconv_loss(xp, pk, Cc, nseg=nseg, argmax=am2)
        sm2 = torch.empty_like(sm)
        ops.deconv_loss(xp, pk, Cc, nseg=nseg, softmax=sm2)
        ok &= bool(torch.equal(am2, am)) and bool(torch.equal(sm2, sm))
    # backward of the stage from a dense output gradient g
    g32 = torch.randn(N, H, W, Cc, device=dev)
    dzp = ops.padded_alloc(N, h, w, s, dev)
    gi = ops.padded_interior(dzp, s)
    gi[0][..., :Cc] = g32.to(torch.bfloat16)
    gi[1][..., :Cc] = (g32 - gi[0][..., :Cc].float()).to(torch.bfloat16)
    gq = ops.planes_to_float(gi, Cc).double() if nseg == 3 else gi[0][..., :Cc].double()
    xr = xq.clone().requires_grad_(True)
    Tr = Tq.clone().requires_grad_(True)
    ref_deconv(xr, Tr, bias, s).backward(gq)
    nxt = ops.planes_alloc(N, h + 2, w + 2, dev, 64)          # dx lands in the interior of a padded tensor
    dxv = tuple(t[:, 1:-1, 1:-1, :] for t in nxt)
    cs = torch.zeros(Cc, device=dev)
    ops.deconv_dx(dzp, pk, Cc, s, dxv, nseg=nseg, colsum=cs)
    got = ops.planes_to_float(dxv, Cc)
    ok &= report("deconv dx " + tag, got, xr.grad, tol)
    ok &= report("deconv dx colsum", cs, got.double().sum((0, 1, 2)), 1e-4)
    dT = torch.full_like(T, float("nan"))
    ops.deconv_dw(xp, dzp, Cc, s, dT, nseg=nseg)
    ok &= report("deconv dT " + tag, dT, Tr.grad, tol)
    return ok


@case
def deconv_stride2():
    ok = deconv_case(20, 2, 2, 6, 10, 3, 2e-5)
    ok &= deconv_case(5, 2, 3, 9, 13, 3, 2e-5)      # ragged tiles of blocks
    ok &= deconv_case(2, 2, 1, 16, 32, 1, 1e-2)     # bf16 mode (hi planes only)
    return ok


@case
def deconv_stride8_loss():
    ok = deconv_case(20, 8, 2, 6, 10, 3, 2e-5)
    ok &= deconv_case(5, 8, 1, 9, 13, 3, 2e-5)      # C % 4 != 0: byte label path
    ok &= deconv_case(2, 8, 2, 12, 8, 3, 2e-5)
    ok &= deconv_case(20, 8, 1, 16, 32, 1, 1e-2)    # bf16 mode; 561 blocks = 5 M tiles (odd: pair partner past the end)
    return ok


@case
def fused_pool():
    """EPI_POOL: the 2x2 max-pool emitted by the conv epilogue equals the stand-alone pool of the stored tensor bit for
    bit, for the halo kernels, the per-tap CTA-pair kernel, bf16 and hi/lo pair storage, with and without the
    full-resolution store."""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(61)
    ok = True
    for (N, H, W, cin, cout, pair) in ((2, 32, 48, 64, 64, False), (1, 20, 36, 64, 128, True), (3, 12, 20, 256, 256, False),
                                        (2, 4, 6, 512, 512, True), (1, 16, 8, 128, 128, False)):
        w = torch.randn(3, 3, cin, cout, device=dev) / (9 * cin) ** 0.5
        b = torch.randn(cout, device=dev)
        wh, wl = _shadow(w, pair)
        x32 = torch.randn(N, H, W, cin, device=dev)
        x = ops.to_pair(x32) if pair else x32.to(torch.bfloat16)
        cm = 2 if pair else 1
        # (one split: the fused variant never splits K, and a split-K sum differs in its last bit)
        y_ref = ops.conv_gemm(x, wh, cout, 3, bias=b, flags=ops.EPI_BIAS | ops.EPI_RELU, wp_lo=wl, pair=pair, w_mode=1,
                              force_splits=1)
        p_ref = ops.maxpool_fwd(y_ref, pair=pair)
        pooled = torch.full((N, H // 2, W // 2, cm * cout), float("nan"), dtype=torch.bfloat16, device=dev)
        y = ops.conv_gemm(x, wh, cout, 3, bias=b, flags=ops.EPI_BIAS | ops.EPI_RELU, wp_lo=wl, pair=pair, w_mode=1,
                          pool_out=pooled)
        pooled2 = torch.full_like(pooled, float("nan"))
        ops.conv_gemm(x, wh, cout, 3, bias=b, flags=ops.EPI_BIAS | ops.EPI_RELU, wp_lo=wl, pair=pair, w_mode=1,
                      pool_out=pooled2, store_out=False)
        torch.cuda.synchronize()
        # the tile shape differs from the unfused run's, the accumulation order inside a tile does not
        val = (lambda t: ops.from_pair(t)) if pair else (lambda t: t.float())      # (+0 / -0 low halves compare equal)
        same = bool(torch.equal(val(y), val(y_ref))) and bool(torch.equal(val(pooled), val(p_ref))) and \
            bool(torch.equal(val(pooled2), val(p_ref)))
        print("  fused pool N%d %dx%d Cin%d Cout%d pair%d: %s" % (N, H, W, cin, cout, pair, "OK" if same else "FAIL"))
        if not same:
            a = ops.from_pair(pooled) if pair else pooled
            r = ops.from_pair(p_ref) if pair else p_ref
            report("   pooled vs stand-alone pool", a, r, 0.0)
            report("   full-res vs unfused", ops.from_pair(y) if pair else y, ops.from_pair(y_ref) if pair else y_ref, 0.0)
        ok &= same
    return ok


@case
def reduced_products():
    """nseg = 2 / 1 on pair operands (the measured reduced-backward modes): x * (w_hi + w_lo) and x_hi * w_hi."""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(62)
    ok = True
    N, H, W, cin, cout, k = 2, 16, 32, 128, 128, 3
    w = torch.randn(k, k, cin, cout, device=dev) / (k * k * cin) ** 0.5
    wh, wl = _shadow(w, True)
    dy = ops.to_pair(torch.randn(N, H, W, cout, device=dev))
    dyh = dy[..., :cout].double()
    for nseg, wq in ((2, wh.double() + wl.double()), (1, wh.double())):
        dx = ops.conv_gemm(dy, wh, cin, k, wp_lo=wl, pair=True, w_mode=2, nseg=nseg)
        ref = F.conv_transpose2d(dyh.permute(0, 3, 1, 2), wq.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
        ok &= report("dgrad nseg %d" % nseg, ops.from_pair(dx), ref, 2e-5)
    x = ops.to_pair(torch.randn(N, H, W, cin, device=dev))
    xh = x[..., :cin].double()
    for nseg, dq in ((2, ops.from_pair(dy).double()), (1, dyh)):
        dw = torch.empty(k * k * cin, cout, device=dev)
        ops.wgrad_gemm(x, dy, k, dw, pair=True, nseg=nseg)
        cols = F.unfold(F.pad(xh.permute(0, 3, 1, 2), (1,) * 4), k).view(N, cin, k * k, H * W)
        ref = torch.einsum("nctp,npo->tco", cols, dq.reshape(N, H * W, cout)).reshape(k * k * cin, cout)
        ok &= report("wgrad nseg %d" % nseg, dw, ref, 2e-5)
    return ok


@case
def conv1_direct():
    """conv1_1 from the uint8 image (im2col operand built in shared memory): forward (bias + ReLU) and the filter
    gradient, bf16 and hi/lo pair, against fp64 references; borders, several images, N*H*W = k * 128."""
    from fcn8s_tensorflow_b200 import ops
    dev = torch.device("cuda")
    torch.manual_seed(71)
    ok = True
    mean = torch.tensor([103.939, 116.779, 123.68], device=dev, dtype=torch.float64)
    for (N, H, W) in ((1, 32, 32), (2, 64, 96), (3, 32, 160)):
        img = torch.randint(0, 256, (N, H, W, 3), dtype=torch.uint8, device=dev)
        bgr = img.double().flip(-1) - mean
        w = torch.randn(3, 3, 3, 64, device=dev) * 0.1
        b = torch.randn(64, device=dev)
        for pair, tol in ((True, 2e-5), (False, 1e-2)):
            pk = ops.pack_weights(w, 1, 27, 64, 0, ops.BF16, cin_pad=64, split=pair)
            wq = (pk[0].double() + (pk[1].double() if pair else 0))[:, :27].t().reshape(3, 3, 3, 64)
            xq = _pair_q(bgr.float()) if pair else bgr.float().to(torch.bfloat16).double()
            out = torch.full((N, H, W, 128 if pair else 64), float("nan"), dtype=torch.bfloat16, device=dev)
            ops.conv1_fwd(img, pk, b, out, pair=pair)
            ref = ref_conv(xq, wq, b, relu=True)
            tag = "N%d %dx%d pair%d" % (N, H, W, pair)
            ok &= report("conv1 fwd " + tag, ops.from_pair(out) if pair else out, ref, tol)
            dy32 = torch.randn(N, H, W, 64, device=dev)
            dy = ops.to_pair(dy32) if pair else dy32.to(torch.bfloat16)
            dyq = ops.from_pair(dy).double() if pair else dy.double()
            dw = torch.full((27, 64), float("nan"), device=dev)
            ops.conv1_wgrad(img, dy, dw, pair=pair)
            cols = F.unfold(F.pad(xq.permute(0, 3, 1, 2), (1,) * 4), 3).view(N, 3, 9, H * W)   # [n, c, tap, p]
            refw = torch.einsum("nctp,npo->tco", cols, dyq.reshape(N, H * W, 64)).reshape(27, 64)
            ok &= report("conv1 wgrad " + tag, dw, refw, tol)
    return ok


@case
def promoted_accumulation():
    """The TMEM accumulator truncates (rounds toward zero) on every tcgen05.mma: with all-positive operands -- the worst
    case, every partial sum has the same sign -- an accumulation of n MMAs falls short of the exact sum by ~n * 7e-8
    relative.  The PROMO kernels cut the hi*hi segment into chunks of promo_kb k-blocks, each accumulated from zero and
    added to an fp32 register sum with round-to-nearest adds (ConvGemmArgs::promo_kb), which bounds the loss by the
    chunk length.  Measured here: mean signed relative error without promotion (fcn8_debug_set(9, -1)), with the
    library default, and with 3-k-block chunks; plus mixed-sign cases against fp64 with many short chunks (chunk
    bookkeeping of the MMA issuer / epilogue warps, pair and single-CTA kernels, split-K partials, ragged tiles)."""
    from fcn8s_tensorflow_b200 import _capi as capi
    from fcn8s_tensorflow_b200 import ops
    lib = capi.load()
    dev = torch.device("cuda")
    torch.manual_seed(31)
    ok = True
    for (n, h, w_, cin, k, cout, force_bn, force_splits) in ((2, 16, 32, 4096, 1, 256, 0, 1), (3, 20, 36, 512, 3, 512, 0, 0),
                                                           (2, 16, 32, 1024, 1, 128, 128, 1), (2, 16, 32, 1024, 1, 64, 64, 1),
                                                           (4, 4, 8, 512, 7, 256, 0, 0)):
        x32 = torch.rand(n, h, w_, cin, device=dev) + 0.5
        wt = (torch.rand(k, k, cin, cout, device=dev) + 0.5) / (k * k * cin)
        wh, wl = _shadow(wt, True)
        x = ops.to_pair(x32)
        ref = ref_conv(ops.from_pair(x).double(), wh.double() + wl.double())
        errs = []
        for P in (-1, 0, 3):
            lib.fcn8_debug_set(9, P)
            try:
                y = ops.conv_gemm(x, wh, cout, k, wp_lo=wl, pair=True, w_mode=1, force_bn=force_bn,
                                  force_splits=force_splits, algo=1)
                torch.cuda.synchronize()
            finally:
                lib.fcn8_debug_set(9, 0)
            errs.append((ops.from_pair(y).double() / ref - 1).mean().item())
        n_mma = k * k * cin // 16
        print("  N%d %dx%d Cin%d k%d Cout%d bn%d: %5d hi*hi MMAs: mean signed rel error  unpromoted %+.2e  default %+.2e  "
              "3-block chunks %+.2e" % (n, h, w_, cin, k, cout, force_bn, n_mma, errs[0], errs[1], errs[2]))
        good = abs(errs[1]) <= 5e-6 and abs(errs[2]) <= 2e-6
        if not good:
            print("    FAIL")
        ok &= good
    lib.fcn8_debug_set(9, 3)
    try:
        ok &= hwio_conv_case(3, 20, 36, 256, 256, 3, True, 2e-5, algo=1)      # pair kernel, odd tile count, ragged
        ok &= hwio_conv_case(1, 4, 8, 512, 256, 7, True, 2e-5)                # split-K partials of promoted tiles
        ok &= hwio_conv_case(2, 16, 32, 256, 128, 3, True, 2e-5, force_bn=128)
        ok &= hwio_conv_case(2, 16, 32, 256, 64, 3, True, 2e-5, force_bn=64)
        ok &= hwio_conv_case(2, 16, 32, 256, 256, 3, True, 2e-5, force_bn=256, algo=1)
    finally:
        lib.fcn8_debug_set(9, 0)
    return ok


@case
def halo_pair_packed_forward():
    """conv_halo_pair_kernel with the packed K-major forward operand (w_mode 0; conv1_2's forward in the fp32-equivalent
    mode: the 32 output channels per CTA of a pair are too narrow for the MN-major operand of the TF layout): hi/lo pair
    and bf16 operands, Cin = 64 (resident half-weights) and 128 (two channel blocks), ragged tiles, odd tile counts,
    with and without the fused max-pool -- against fp64 and against the TF-layout path (w_mode 1)."""
    from fcn8s_tensorflow_b200 import ops
    torch.manual_seed(17)
    dev = torch.device("cuda")
    ok = True
    for (n, h, w_, cin, pair) in ((2, 32, 48, 64, True), (3, 20, 36, 64, True), (1, 16, 8, 128, True), (2, 32, 48, 64, False),
                                  (1, 40, 24, 128, False)):
        w = torch.randn(3, 3, cin, 64, device=dev) / (9 * cin) ** 0.5
        b = torch.randn(64, device=dev)
        wp, wpl = ops.pack_weights(w, 3, cin, 64, 0, ops.BF16, split=pair)
        wh, wl = _shadow(w, pair)
        wq = (wp.double() + wpl.double()) if pair else wp.double()
        wq = wq.view(64, 9, cin).permute(1, 2, 0).reshape(3, 3, cin, 64)
        x32 = torch.randn(n, h, w_, cin, device=dev)
        x = ops.to_pair(x32) if pair else x32.to(torch.bfloat16)
        xq = ops.from_pair(x).double() if pair else x.double()
        fl = ops.EPI_BIAS | ops.EPI_RELU
        y = ops.conv_gemm(x, wp, 64, 3, bias=b, flags=fl, wp_lo=wpl, pair=pair, w_mode=0, algo=2)
        y_tf = ops.conv_gemm(x, wh, 64, 3, bias=b, flags=fl, wp_lo=wl, pair=pair, w_mode=1, algo=2)
        torch.cuda.synchronize()
        got = ops.from_pair(y) if pair else y
        tag = "N%d %dx%d Cin%d pair%d" % (n, h, w_, cin, pair)
        ok &= report("packed fwd " + tag, got, ref_conv(xq, wq, b, relu=True), 2e-5 if pair else 1e-2)
        ok &= report("packed == TF-layout path " + tag, got, (ops.from_pair(y_tf) if pair else y_tf).double(),
                     1e-6 if pair else 1e-2)
        if h % 2 == 0 and w_ % 2 == 0:
            cm = 2 if pair else 1
            pooled = torch.full((n, h // 2, w_ // 2, cm * 64), float("nan"), dtype=torch.bfloat16, device=dev)
            y2 = ops.conv_gemm(x, wp, 64, 3, bias=b, flags=fl, wp_lo=wpl, pair=pair, w_mode=0, algo=2, pool_out=pooled)
            p_ref = ops.maxpool_fwd(y, pair=pair)
            torch.cuda.synchronize()
            val = (lambda t: ops.from_pair(t)) if pair else (lambda t: t.float())
            same = bool(torch.equal(val(y2), val(y))) and bool(torch.equal(val(pooled), val(p_ref)))
            print("  %-58s %s" % ("fused pool " + tag, "OK" if same else "FAIL"))
            ok &= same
    return ok


@case
def dynamic_tiles():
    """Dynamic tile scheduling (cluster launch control, fcn8_debug_set(10, 1)): the grid holds one CTA / CTA pair per
    tile, running CTAs cancel pending ones and take their tiles.  Same references as the static cases -- per-tap, pair,
    promoted, split-K and halo forward / dgrad kernels, the filter-gradient kernels -- plus bit-for-bit equality with
    the static schedule (a tile's arithmetic does not depend on which CTA runs it).  scripts/clc_probe.cu is the
    stand-alone probe of the mechanism."""
    from fcn8s_tensorflow_b200 import _capi, ops
    lib = _capi.load()
    torch.manual_seed(41)
    dev = torch.device("cuda")
    w = torch.randn(3, 3, 256, 256, device=dev) / 48.0
    wh, wl = _shadow(w, True)
    x = ops.to_pair(torch.randn(5, 44, 52, 256, device=dev))      # 90 M tiles: more than one wave of CTA pairs
    w64 = torch.randn(3, 3, 64, 64, device=dev) / 24.0
    wh64, _ = _shadow(w64, False)
    x64 = torch.randn(3, 100, 120, 64, device=dev).to(torch.bfloat16)   # 315 halo tiles
    y_static = ops.conv_gemm(x, wh, 256, 3, wp_lo=wl, pair=True, w_mode=1, algo=1)
    h_static = ops.conv_gemm(x64, wh64, 64, 3, w_mode=1)
    lib.fcn8_debug_set(10, 1)
    try:
        ok = hwio_conv_case(3, 20, 36, 256, 256, 3, True, 2e-5, algo=1)       # pair + promoted, odd tile count
        ok &= hwio_conv_case(1, 4, 8, 512, 256, 7, True, 2e-5)                # split-K partials
        ok &= hwio_conv_case(2, 16, 32, 256, 128, 3, False, 1e-2, force_bn=128)
        ok &= hwio_conv_case(3, 20, 36, 64, 128, 3, True, 2e-5)               # halo kernels, ragged tiles
        ok &= hwio_conv_case(2, 16, 32, 64, 64, 3, False, 1e-2)               # halo, resident weights
        ok &= wgrad_case(3, 20, 36, 128, 256, 3, 0, 1e-2)
        ok &= wgrad_case(2, 16, 32, 256, 512, 3, 0, 1e-2)
        ok &= wgrad_case(1, 8, 16, 512, 256, 7, 0, 1e-2)
        y_dyn = ops.conv_gemm(x, wh, 256, 3, wp_lo=wl, pair=True, w_mode=1, algo=1)
        h_dyn = ops.conv_gemm(x64, wh64, 64, 3, w_mode=1)
        torch.cuda.synchronize()
    finally:
        lib.fcn8_debug_set(10, 0)
    same = bool(torch.equal(y_dyn, y_static)) and bool(torch.equal(h_dyn, h_static))
    print("  %-58s %s" % ("dynamic schedule == static schedule, bit for bit", "OK" if same else "FAIL"))
    return ok and same


def main():
    args = sys.argv[1:]
    if not args or args[0] == "list":
        print("\n".join(CASES))
        return 0
    if args[0] == "all":
        results = {}
        per_case_timeout = int(os.environ.get("BRINGUP_TIMEOUT", "240"))
        for name in CASES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=per_case_timeout,
                                   stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                out, code = r.stdout, r.returncode
            except subprocess.TimeoutExpired as e:
                out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
                code = -999
            results[name] = code
            print("==== %s: exit %d (%.1fs)" % (name, code, time.time() - t0))
            print(out[-6000:])
            sys.stdout.flush()
        print(json.dumps(results))
        return 0 if all(v == 0 for v in results.values()) else 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ok = True
    for name in args:
        print("case %s" % name)
        try:
            r = CASES[name]()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            r = False
        print("case %s -> %s" % (name, "PASS" if r else "FAIL"))
        ok &= bool(r)
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
