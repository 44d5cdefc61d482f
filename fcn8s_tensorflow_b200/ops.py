"""Operator layer: one Python function per C-ABI entry point, taking/returning torch CUDA tensors.

torch is used for device memory and the current stream only; every computation goes through
libfcn8s_sm100.so. Shapes follow the reference's NHWC convention (fcn8s_tensorflow.py:429-433).
"""
import ctypes as C

import torch

from . import _capi as capi
from ._capi import (BF16, F32, BF16X2, EPI_BIAS, EPI_RELU, EPI_DROPOUT, EPI_MASK, EPI_RESIDUAL,  # noqa: F401
                    EPI_ROUND_TF32, EPI_COLSUM)

_workspaces = {}


class KernelTimer:
    """Optional per-launch CUDA-event timing of the tensor-core GEMM kernels on the launching stream (bench.py's
    roofline numbers). Install with `ops.TIMER = KernelTimer()`; records (tag, algorithmic flops, start, end)."""

    def __init__(self):
        self.records = []
        self._pool = []

    def _event(self):
        return self._pool.pop() if self._pool else torch.cuda.Event(enable_timing=True)

    def start(self):
        e = self._event()
        e.record(torch.cuda.current_stream())
        return e

    def stop(self, tag, flops, e0):
        e1 = self._event()
        e1.record(torch.cuda.current_stream())
        self.records.append((tag, flops, e0, e1))

    def summary(self):
        """tag -> dict(launches, flops, ms); call after a synchronize. Clears the records."""
        out = {}
        for tag, flops, e0, e1 in self.records:
            d = out.setdefault(tag, dict(launches=0, flops=0.0, ms=0.0))
            d["launches"] += 1
            d["flops"] += flops
            d["ms"] += e0.elapsed_time(e1)
            self._pool += [e0, e1]
        self.records = []
        return out


TIMER = None


def torch_dtype(dtype):
    return torch.float32 if dtype == F32 else torch.bfloat16


def to_pair(x):
    """fp32 [..., C] -> bf16 hi/lo pair [..., 2C] (FCN8_BF16X2): hi = bf16(x), lo = bf16(x - hi). Test helper."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=-1).contiguous()


def from_pair(xp):
    """bf16 hi/lo pair [..., 2C] -> fp32 [..., C]."""
    Cc = xp.shape[-1] // 2
    return xp[..., :Cc].float() + xp[..., Cc:].float()


def dtype_of(t):
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise TypeError("activation tensors must be bfloat16 or float32, got %s" % t.dtype)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _workspace(nbytes, device):
    """Grow-only scratch buffer per device (caller-provided workspace of the C ABI)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _chk_cuda(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise capi.Fcn8Error("tensor is not on a CUDA device; there is no CPU path")
            if not t.is_contiguous():
                raise capi.Fcn8Error("tensor must be contiguous")


def preprocess_im2col(images, dtype):
    """uint8 RGB [N,H,W,3] -> mean-subtracted BGR im2col [N,H,W,KP] for conv1_1 (KP = 64 bf16 / 32 f32)."""
    _chk_cuda(images)
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[3] != 3:
        raise ValueError("images must be uint8 [N,H,W,3]")
    N, H, W, _ = images.shape
    kp = {BF16: 64, F32: 32, BF16X2: 128}[dtype]   # BF16X2: 64 hi + 64 lo columns
    out = torch.empty((N, H, W, kp), dtype=torch_dtype(dtype), device=images.device)
    p = capi.PreprocessParams(capi.ptr(images), capi.ptr(out), N, H, W, dtype)
    capi.check(capi.load().fcn8_preprocess_im2col(C.byref(p), _stream()))
    return out


def pack_weights(w, ksize, cin, cout, mode, dtype, cin_pad=None, split=False, out=None):
    """fp32 HWIO weights -> tensor-core operand layout (mode 0 fprop [Cout][taps*CinPad], mode 1 dgrad
    [Cin][taps*Cout]). Returns (packed, packed_lo or None); `out` = an earlier result to refill in place."""
    _chk_cuda(w)
    taps = ksize * ksize
    if cin_pad is None:
        cin_pad = cin
    shape = (cout, taps * cin_pad) if mode == 0 else (cin, taps * cout)
    if out is not None:
        out, lo = out
    else:
        out = torch.empty(shape, dtype=torch_dtype(dtype), device=w.device)
        lo = torch.empty_like(out) if split else None
    p = capi.PackParams(capi.ptr(w), capi.ptr(out), capi.ptr(lo), ksize, cin, cout, cin_pad, mode, dtype)
    capi.check(capi.load().fcn8_pack_weights(C.byref(p), _stream()))
    return out, lo


def split_tf32(x):
    """(trunc_tf32(x), round_tf32(x - trunc_tf32(x)))"""
    _chk_cuda(x)
    hi, lo = torch.empty_like(x), torch.empty_like(x)
    capi.check(capi.load().fcn8_split_tf32(capi.ptr(x), capi.ptr(hi), capi.ptr(lo), x.numel(), _stream()))
    return hi, lo


def conv_gemm(x, wp, cout, ksize, bias=None, flags=0, mask_src=None, residual=None, mask_scale=1.0, keep_prob=1.0,
              seed=0, out=None, x_lo=None, wp_lo=None, force_splits=0, force_bn=0, pair=False, w_mode=0, colsum=None,
              seed_ptr=None, algo=0, nseg=3):
    """Stride-1 SAME k x k convolution (fprop or dgrad, see include/fcn8s_b200.h).
    pair=True: x / out / mask_src / residual are bf16 hi/lo pair tensors [N,H,W,2C] (FCN8_BF16X2) and the product is
    the error-compensated hi*hi + hi*lo + lo*hi (needs wp_lo); nseg = 2 / 1 keep only the first two / one of those
    products (the measured reduced-backward modes).  w_mode 1 / 2: wp (wp_lo) is the bf16 shadow of the TF weight tensor
    itself (fprop / dgrad), no packing."""
    _chk_cuda(x, wp, bias, mask_src, residual, out, x_lo, wp_lo, colsum)
    if colsum is not None:
        flags |= EPI_COLSUM
    N, H, W, cin = x.shape
    if pair:
        cin //= 2
        dtype = BF16
        if out is None:
            out = torch.empty((N, H, W, 2 * cout), dtype=torch.bfloat16, device=x.device)
        p = capi.ConvParams(capi.ptr(x), capi.ptr(x, cin), capi.ptr(wp), capi.ptr(wp_lo), capi.ptr(out), capi.ptr(bias),
                            capi.ptr(mask_src), capi.ptr(residual), N, H, W, cin, cout, ksize, dtype, nseg, flags,
                            mask_scale, keep_prob, seed, force_splits, force_bn, 2 * cin, 2 * cout,
                            capi.ptr(out, cout), capi.ptr(residual, cout), w_mode, capi.ptr(colsum),
                            capi.ptr(seed_ptr), algo)
    else:
        dtype = dtype_of(x)
        if out is None:
            out = torch.empty((N, H, W, cout), dtype=x.dtype, device=x.device)
        nseg = 3 if x_lo is not None else 1
        p = capi.ConvParams(capi.ptr(x), capi.ptr(x_lo), capi.ptr(wp), capi.ptr(wp_lo), capi.ptr(out), capi.ptr(bias),
                            capi.ptr(mask_src), capi.ptr(residual), N, H, W, cin, cout, ksize, dtype, nseg, flags,
                            mask_scale, keep_prob, seed, force_splits, force_bn, 0, 0, None, None, w_mode,
                            capi.ptr(colsum), capi.ptr(seed_ptr), algo)
    lib = capi.load()
    nbytes = lib.fcn8_conv_gemm_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device) if nbytes else None
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(lib.fcn8_conv_gemm(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    if e0 is not None:
        TIMER.stop("conv_gemm", 2.0 * N * H * W * cout * ksize * ksize * cin, e0)
    return out


def wgrad_gemm(x, dy, ksize, out, rows_valid=0, x_lo=None, dy_lo=None, force_splits=0, force_bn=0, pair=False,
               nseg=3):
    """Filter gradient into `out` (fp32, HWIO-flattened [k*k*Cin, Cout] or its first rows_valid rows).
    pair=True: x / dy are bf16 hi/lo pair tensors [N,H,W,2C]; error-compensated product."""
    _chk_cuda(x, dy, out, x_lo, dy_lo)
    N, H, W, cin = x.shape
    cout = dy.shape[3]
    if pair:
        cin //= 2
        cout //= 2
        p = capi.WgradParams(capi.ptr(x), capi.ptr(x, cin), capi.ptr(dy), capi.ptr(dy, cout), capi.ptr(out), N, H, W,
                             cin, cout, ksize, rows_valid, BF16, nseg, force_splits, force_bn, 2 * cin, 2 * cout)
    else:
        nseg = 3 if x_lo is not None else 1
        p = capi.WgradParams(capi.ptr(x), capi.ptr(x_lo), capi.ptr(dy), capi.ptr(dy_lo), capi.ptr(out), N, H, W, cin,
                             cout, ksize, rows_valid, dtype_of(x), nseg, force_splits, force_bn, 0, 0)
    lib = capi.load()
    nbytes = lib.fcn8_wgrad_gemm_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device) if nbytes else None
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(lib.fcn8_wgrad_gemm(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    if e0 is not None:
        TIMER.stop("wgrad_gemm", 2.0 * N * H * W * cout * ksize * ksize * cin, e0)
    return out


def _fmt(t, pair):
    return BF16X2 if pair else dtype_of(t)


def maxpool_fwd(x, out=None, pair=False):
    _chk_cuda(x, out)
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((N, (H + 1) // 2, (W + 1) // 2, Cc), dtype=x.dtype, device=x.device)
    p = capi.PoolParams(capi.ptr(x), capi.ptr(out), None, N, H, W, Cc // 2 if pair else Cc, _fmt(x, pair), None)
    capi.check(capi.load().fcn8_maxpool_fwd(C.byref(p), _stream()))
    return out


def maxpool_bwd(x, dy, out=None, pair=False, db=None):
    """Gradient w.r.t. the pre-ReLU producer of x: routes dy to the first arg-max where x > 0.
    db (fp32 [C], accumulated): bias gradient of that producer (sum of the result over pixels)."""
    _chk_cuda(x, dy, out, db)
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty_like(x)
    p = capi.PoolParams(capi.ptr(x), capi.ptr(dy), capi.ptr(out), N, H, W, Cc // 2 if pair else Cc, _fmt(x, pair),
                        capi.ptr(db))
    capi.check(capi.load().fcn8_maxpool_bwd(C.byref(p), _stream()))
    return out


def bias_grad(dy, out, pair=False):
    _chk_cuda(dy, out)
    Cc = dy.shape[-1]
    P = dy.numel() // Cc
    p = capi.BiasGradParams(capi.ptr(dy), capi.ptr(out), P, Cc // 2 if pair else Cc, _fmt(dy, pair))
    lib = capi.load()
    nbytes = lib.fcn8_bias_grad_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, dy.device)
    capi.check(lib.fcn8_bias_grad(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    return out


def score_head_fwd(x, K, b, scale, out=None, pair=False):
    _chk_cuda(x, K, b, out)
    cin = x.shape[-1]
    Cc = K.shape[-1]
    P = x.numel() // cin
    if pair:
        cin //= 2
    if out is None:
        out = torch.empty(tuple(x.shape[:-1]) + (Cc,), dtype=torch.float32, device=x.device)
    p = capi.HeadParams(capi.ptr(x), capi.ptr(K), capi.ptr(b), capi.ptr(out), None, None, None, P, cin, Cc, scale,
                        _fmt(x, pair), 0, 1.0)
    lib = capi.load()
    nbytes = lib.fcn8_score_head_fwd_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device) if nbytes else None
    capi.check(lib.fcn8_score_head_fwd(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    return out


def score_head_bwd(x, K, ds, scale, dK, db, dx=None, mask=False, mask_scale=1.0, pair=False):
    _chk_cuda(x, K, ds, dK, db, dx)
    cin = x.shape[-1]
    Cc = K.shape[-1]
    P = x.numel() // cin
    if pair:
        cin //= 2
    p = capi.HeadParams(capi.ptr(x), capi.ptr(K), None, capi.ptr(ds), capi.ptr(dK), capi.ptr(db), capi.ptr(dx), P, cin,
                        Cc, scale, _fmt(x, pair), 1 if mask else 0, mask_scale)
    lib = capi.load()
    nbytes = lib.fcn8_score_head_bwd_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device)
    capi.check(lib.fcn8_score_head_bwd(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    return dx


def upscore_fwd(x, T, bias, stride, skip=None, out=None):
    _chk_cuda(x, T, bias, skip, out)
    N, h, w, Cc = x.shape
    if out is None:
        out = torch.empty((N, h * stride, w * stride, Cc), dtype=torch.float32, device=x.device)
    p = capi.UpscoreParams(capi.ptr(x), capi.ptr(T), capi.ptr(bias), capi.ptr(skip), capi.ptr(out), None, None, None,
                           N, h, w, Cc, stride)
    capi.check(capi.load().fcn8_upscore_fwd(C.byref(p), _stream()))
    return out


def upscore_bwd(x, T, dy, stride, dT, dbias, dx=None):
    _chk_cuda(x, T, dy, dT, dbias, dx)
    N, h, w, Cc = x.shape
    p = capi.UpscoreParams(capi.ptr(x), capi.ptr(T), None, None, capi.ptr(dy), capi.ptr(dx), capi.ptr(dT),
                           capi.ptr(dbias), N, h, w, Cc, stride)
    lib = capi.load()
    nbytes = lib.fcn8_upscore_bwd_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device)
    capi.check(lib.fcn8_upscore_bwd(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    return dx


def softmax_xent(logits, labels=None, loss_sum=None, dlogits=None, softmax=None, argmax=None, grad_scale=1.0,
                 dbias=None, pad=0, num_classes=None):
    """Loss / predictor kernel. `logits` (and `dlogits`) are [N,Hp,Wp,CP] with Hp = H + 2*pad, Wp = W + 2*pad and
    CP >= num_classes (pad = 0, CP = C: the dense tensor); labels / softmax / argmax are dense over [N,H,W]."""
    _chk_cuda(logits, labels, loss_sum, dlogits, softmax, argmax, dbias)
    if logits.dim() != 4:
        logits = logits.view(1, 1, -1, logits.shape[-1])
    N, Hp, Wp, CP = logits.shape
    Cc = CP if num_classes is None else num_classes
    p = capi.SoftmaxParams(capi.ptr(logits), capi.ptr(labels), capi.ptr(loss_sum), capi.ptr(dlogits),
                           capi.ptr(dbias), capi.ptr(softmax), capi.ptr(argmax), N, Hp - 2 * pad, Wp - 2 * pad, Cc,
                           CP, pad, grad_scale)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(capi.load().fcn8_softmax_xent(C.byref(p), _stream()))
    if e0 is not None:   # HBM-bound: the "work" recorded is algorithmic bytes, not FLOPs
        px = N * (Hp - 2 * pad) * (Wp - 2 * pad)
        nbytes = px * Cc * 4 + (px * Cc if labels is not None else 0) + (px * Cc * 4 if dlogits is not None else 0) + \
            (px * Cc * 4 if softmax is not None else 0) + (px * 8 if argmax is not None else 0)
        TIMER.stop("predictor" if dlogits is None else "loss", float(nbytes), e0)


def upscore_tc_cp(num_classes, stride):
    """Channel stride of the padded blocked output of the tensor-core transposed convolution."""
    return capi.load().fcn8_upscore_tc_cp(num_classes, stride)


def upscore_tc_pack(T, bias, stride, split=False, out=None):
    """T [2s,2s,C,C], bias [C] -> dict of tensor-core operands (w_fwd, w_dx, bias_big and their *_lo halves).
    `out`: a dict returned by an earlier call, refilled in place."""
    _chk_cuda(T, bias)
    Cc = T.shape[-1]
    CP = upscore_tc_cp(Cc, stride)
    ncols = stride * stride * CP
    f32 = dict(dtype=torch.float32, device=T.device)
    if out is None:
        out = dict(w_fwd=torch.empty((ncols, 128), **f32), w_dx=torch.empty((64, 4 * ncols), **f32),
                   bias_big=torch.empty(ncols, **f32), w_fwd_lo=None, w_dx_lo=None)
        if split:
            out["w_fwd_lo"] = torch.empty_like(out["w_fwd"])
            out["w_dx_lo"] = torch.empty_like(out["w_dx"])
    p = capi.UpscorePackParams(capi.ptr(T), capi.ptr(bias), capi.ptr(out["w_fwd"]), capi.ptr(out["w_fwd_lo"]),
                               capi.ptr(out["w_dx"]), capi.ptr(out["w_dx_lo"]), capi.ptr(out["bias_big"]), Cc, stride)
    capi.check(capi.load().fcn8_upscore_tc_pack(C.byref(p), _stream()))
    return out


def _tc_params(x, x_lo, w, w_lo, bias_big, zp, zp_lo, dx, dT, Cc, stride):
    N, h, wd, ldx = x.shape if x is not None else dx.shape
    nseg = 3 if (w_lo is not None or (x_lo is not None and zp_lo is not None)) else 1
    return capi.UpscoreTcParams(capi.ptr(x), capi.ptr(x_lo), capi.ptr(w), capi.ptr(w_lo), capi.ptr(bias_big),
                                capi.ptr(zp), capi.ptr(zp_lo), capi.ptr(dx), capi.ptr(dT), N, h, wd, Cc, stride, ldx,
                                nseg)


def upscore_tc_alloc(N, h, w, num_classes, stride, device, zero=False):
    """Padded blocked tensor [N, s*(h+1), s*(w+1), CP] of the tensor-core transposed convolution."""
    CP = upscore_tc_cp(num_classes, stride)
    shape = (N, stride * (h + 1), stride * (w + 1), CP)
    return (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=device)


def upscore_tc_interior(zp, num_classes, stride):
    """[N, s*h, s*w, C] view of the transposed convolution's output inside its padded tensor."""
    p = stride // 2
    return zp[:, p:zp.shape[1] - p, p:zp.shape[2] - p, :num_classes]


def upscore_tc_gather(zp, skip, out, num_classes, stride):
    """out[N, s*h, s*w, ld] = interior(zp)[..., :C] (+ skip [N, s*h, s*w, >=C]); pad channels of `out` zeroed."""
    _chk_cuda(zp, skip, out)
    N, H, W, ld = out.shape
    capi.check(capi.load().fcn8_upscore_tc_gather(capi.ptr(zp), capi.ptr(skip), capi.ptr(out), N, H // stride,
                                                  W // stride, num_classes, stride, ld,
                                                  skip.shape[-1] if skip is not None else 0, _stream()))
    return out


def upscore_tc_scatter(g, dzp, num_classes, stride, dbias=None):
    """dzp interior <- g [N, s*h, s*w, ld]; dbias[c] += sum of g over pixels (caller zeroes dzp once and dbias)."""
    _chk_cuda(g, dzp, dbias)
    N, H, W, ld = g.shape
    capi.check(capi.load().fcn8_upscore_tc_scatter(capi.ptr(g), capi.ptr(dzp), capi.ptr(dbias), N, H // stride,
                                                   W // stride, num_classes, stride, ld, _stream()))
    return dzp


def upscore_tc_fwd(x, packed, num_classes, stride, out, x_lo=None):
    """x [N,h,w,ldx] fp32 (channels >= C zero) -> padded blocked output `out` (see upscore_tc_alloc)."""
    _chk_cuda(x, x_lo, out)
    w_lo = packed["w_fwd_lo"] if x_lo is not None else None
    p = _tc_params(x, x_lo, packed["w_fwd"], w_lo, packed["bias_big"], out, None, None, None, num_classes, stride)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(capi.load().fcn8_upscore_tc_fwd(C.byref(p), _stream()))
    if e0 is not None:
        N, h, w, _ = x.shape
        TIMER.stop("upscore8" if stride == 8 else "upscore_tc",
                   2.0 * N * h * w * 4 * stride * stride * num_classes * num_classes, e0)
    return out


def upscore_tc_dx(dzp, packed, num_classes, stride, dx, dzp_lo=None):
    """Input gradient of the transposed convolution from the padded blocked dz (zero border) into dx [N,h,w,ldx]."""
    _chk_cuda(dzp, dzp_lo, dx)
    w_lo = packed["w_dx_lo"] if dzp_lo is not None else None
    p = _tc_params(None, None, packed["w_dx"], w_lo, None, dzp, dzp_lo, dx, None, num_classes, stride)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(capi.load().fcn8_upscore_tc_dx(C.byref(p), _stream()))
    if e0 is not None:
        N, h, w, _ = dx.shape
        TIMER.stop("upscore_tc", 2.0 * N * h * w * 4 * stride * stride * num_classes * num_classes, e0)
    return dx


def upscore_tc_dw(x, dzp, num_classes, stride, dT, x_lo=None, dzp_lo=None):
    """Filter gradient dT [2s,2s,C,C] (TF layout) of the transposed convolution."""
    _chk_cuda(x, dzp, dT, x_lo, dzp_lo)
    p = _tc_params(x, x_lo, None, None, None, dzp, dzp_lo, None, dT, num_classes, stride)
    lib = capi.load()
    nbytes = lib.fcn8_upscore_tc_dw_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(lib.fcn8_upscore_tc_dw(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    if e0 is not None:
        N, h, w, _ = x.shape
        TIMER.stop("upscore_tc", 2.0 * N * h * w * 4 * stride * stride * num_classes * num_classes, e0)
    return dT


def confusion_matrix(pred, labels_onehot, conf):
    _chk_cuda(pred, labels_onehot, conf)
    Cc = labels_onehot.shape[-1]
    capi.check(capi.load().fcn8_confusion_matrix(capi.ptr(pred), capi.ptr(labels_onehot), capi.ptr(conf),
                                                 pred.numel(), Cc, _stream()))


def adam(p, g, m, v, lr_t, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0, w_hi=None, w_lo=None, lr_ptr=None,
         g_bf16=None):
    """TF-form Adam over a flat buffer; w_hi / w_lo (bf16, same length): tensor-core shadow refreshed in the pass.
    lr_ptr: device fp32 scalar overriding lr_t (see set_step_scalars).  g_bf16: the gradient as a bf16 buffer (what a
    bf16 all-reduce leaves), read instead of g."""
    _chk_cuda(p, g, m, v, w_hi, w_lo, lr_ptr, g_bf16)
    capi.check(capi.load().fcn8_adam(capi.ptr(p), capi.ptr(g), capi.ptr(m), capi.ptr(v), p.numel(), lr_t, beta1,
                                     beta2, eps, grad_scale, capi.ptr(w_hi), capi.ptr(w_lo), capi.ptr(lr_ptr),
                                     capi.ptr(g_bf16), _stream()))


def cast_bf16(x, out):
    """out (bf16, same length) = bf16(x) for a flat fp32 buffer (wire format of the bf16 gradient all-reduce)."""
    _chk_cuda(x, out)
    capi.check(capi.load().fcn8_cast_bf16(capi.ptr(x), capi.ptr(out), x.numel(), _stream()))
    return out


def set_step_scalars(scalars, lr_t, seed):
    """scalars (fp32 [>=2], device): [0] = lr_t, [1] = dropout seed bits."""
    _chk_cuda(scalars)
    capi.check(capi.load().fcn8_set_step_scalars(capi.ptr(scalars), lr_t, int(seed) & 0xFFFFFFFF, _stream()))


def shadow_weights(p, w_hi, w_lo=None):
    """w_hi = bf16(p), w_lo = bf16(p - w_hi)."""
    _chk_cuda(p, w_hi, w_lo)
    capi.check(capi.load().fcn8_shadow_weights(capi.ptr(p), capi.ptr(w_hi), capi.ptr(w_lo), p.numel(), _stream()))


def l2_reg(w, g, loss_sum, rate):
    _chk_cuda(w, g, loss_sum)
    capi.check(capi.load().fcn8_l2_reg(capi.ptr(w), capi.ptr(g), capi.ptr(loss_sum), w.numel(), rate, _stream()))
