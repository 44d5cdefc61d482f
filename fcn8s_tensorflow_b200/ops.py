"""Operator layer: one Python function per C-ABI entry point, taking/returning torch CUDA tensors.

torch is used for device memory and the current stream only; every computation goes through
libfcn8s_sm100.so. Shapes follow the reference's NHWC convention (fcn8s_tensorflow.py:429-433).
"""
import ctypes as C

import torch

from . import _capi as capi
from ._capi import (BF16, F32, BF16X2, EPI_BIAS, EPI_RELU, EPI_DROPOUT, EPI_MASK, EPI_RESIDUAL,  # noqa: F401
                    EPI_ROUND_TF32, EPI_COLSUM, EPI_POOL)

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


def _chk_view(*ts):
    """4-D NHWC views (possibly strided: the interior of a padded tensor, one half of a hi/lo pair tensor)."""
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise capi.Fcn8Error("tensor is not on a CUDA device; there is no CPU path")
            if t.dim() != 4 or t.stride(3) != 1:
                raise capi.Fcn8Error("expected an NHWC view with contiguous channels")


def _geom(t):
    """(pixel stride, row stride, image stride) in elements of an NHWC view."""
    return t.stride(2), t.stride(1), t.stride(0)


def halves(t, c):
    """(hi, lo) channel views of a bf16 hi/lo pair tensor [..., 2c] (FCN8_BF16X2)."""
    return t[..., :c], t[..., c:2 * c]


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
              seed_ptr=None, algo=0, nseg=None, out_pair=None, out_scale=0.0, colsum_n=0, pool_out=None,
              store_out=True, cin=None, tag="conv_gemm"):
    """Stride-1 SAME k x k convolution (fprop or dgrad, see include/fcn8s_b200.h).
    x: NHWC view (may be strided).  pair=True: x is a bf16 hi/lo pair tensor [N,H,W,2C] (FCN8_BF16X2) -- otherwise an
    explicit `x_lo` view supplies the low plane -- and the product is the error-compensated hi*hi + hi*lo + lo*hi
    (needs wp_lo); nseg = 2 / 1 keep only hi*(w_hi + w_lo) / hi*hi.  out_pair (default: pair): out / mask_src / residual /
    pool_out are pair tensors [.., 2*cout].  w_mode 1 / 2: wp (wp_lo) is a bf16 weight tensor in TF layout read in place
    (fprop / dgrad), no packing.  pool_out: also emit the 2x2 max-pool (EPI_POOL); store_out=False then keeps only it."""
    if pair:
        c = x.shape[3] // 2
        x, x_lo = halves(x, c)
    cin = x.shape[3] if cin is None else cin
    if out_pair is None:
        out_pair = pair
    _chk_view(x, x_lo)
    _chk_cuda(wp, bias, wp_lo, colsum)
    N, H, W, _ = x.shape
    dtype = dtype_of(x)
    if nseg is None:
        nseg = 3 if x_lo is not None else 1
    if out is None and store_out:
        out = torch.empty((N, H, W, (2 if out_pair else 1) * cout), dtype=x.dtype, device=x.device)

    def hl(t):
        if t is None:
            return None, None
        return halves(t, cout) if out_pair else (t, None)
    o_hi, o_lo = hl(out)
    r_hi, r_lo = hl(residual)
    m_hi, _ = hl(mask_src)
    p_hi, p_lo = hl(pool_out)
    ref = o_hi if o_hi is not None else None
    _chk_view(o_hi, r_hi, m_hi, p_hi)
    if colsum is not None:
        flags |= EPI_COLSUM
    if pool_out is not None:
        flags |= EPI_POOL
    x_ld, x_sH, x_sN = _geom(x)
    o_ld, o_sH, o_sN = _geom(ref) if ref is not None else (0, 0, 0)
    for t in (r_hi, m_hi):
        if t is not None and ref is not None and _geom(t) != _geom(ref):
            raise capi.Fcn8Error("mask_src / residual must have the geometry of out")
    p = capi.ConvParams(capi.ptr(x), capi.ptr(x_lo), capi.ptr(wp), capi.ptr(wp_lo), capi.ptr(o_hi), capi.ptr(bias),
                        capi.ptr(m_hi), capi.ptr(r_hi), N, H, W, cin, cout, ksize, dtype, nseg, flags, mask_scale,
                        keep_prob, seed, force_splits, force_bn, x_ld, o_ld, capi.ptr(o_lo), capi.ptr(r_lo), w_mode,
                        capi.ptr(colsum), capi.ptr(seed_ptr), algo, out_scale, colsum_n, capi.ptr(p_hi),
                        capi.ptr(p_lo), p_hi.stride(2) if p_hi is not None else 0, x_sH, x_sN, o_sH, o_sN)
    lib = capi.load()
    nbytes = lib.fcn8_conv_gemm_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device) if nbytes else None
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(lib.fcn8_conv_gemm(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    if e0 is not None:
        TIMER.stop(tag, 2.0 * N * H * W * cout * ksize * ksize * cin, e0)
    return out


def wgrad_gemm(x, dy, ksize, out, rows_valid=0, x_lo=None, dy_lo=None, force_splits=0, force_bn=0, pair=False,
               nseg=None, out_cols=0, out_scale=0.0, dy_pair=None):
    """Filter gradient into `out` (fp32, HWIO-flattened [k*k*Cin, Cout] or its first rows_valid rows; [.., out_cols]
    when out_cols is given).  x / dy: NHWC views; pair=True: x (and dy unless dy_pair=False) are bf16 hi/lo pair tensors
    [N,H,W,2C]; otherwise x_lo / dy_lo supply the low planes.  Error-compensated product for nseg = 3."""
    if dy_pair is None:
        dy_pair = pair
    if pair:
        x, x_lo = halves(x, x.shape[3] // 2)
    if dy_pair:
        dy, dy_lo = halves(dy, dy.shape[3] // 2)
    _chk_view(x, dy, x_lo, dy_lo)
    _chk_cuda(out)
    N, H, W, cin = x.shape
    cout = dy.shape[3]
    if nseg is None:
        nseg = 3 if (x_lo is not None and dy_lo is not None) else 1
    x_ld, x_sH, x_sN = _geom(x)
    d_ld, d_sH, d_sN = _geom(dy)
    p = capi.WgradParams(capi.ptr(x), capi.ptr(x_lo), capi.ptr(dy), capi.ptr(dy_lo), capi.ptr(out), N, H, W, cin, cout,
                         ksize, rows_valid, dtype_of(x), nseg, force_splits, force_bn, x_ld, d_ld, out_cols, out_scale,
                         x_sH, x_sN, d_sH, d_sN)
    lib = capi.load()
    nbytes = lib.fcn8_wgrad_gemm_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x.device) if nbytes else None
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(lib.fcn8_wgrad_gemm(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    if e0 is not None:
        TIMER.stop("wgrad_gemm" if out_cols == 0 else "head_gemm", 2.0 * N * H * W * cout * ksize * ksize * cin, e0)
    return out


def conv1_fwd(images, packed, bias, out, pair=False):
    """conv1_1 + bias + ReLU straight from the uint8 image [N,H,W,3] (im2col operand built in shared memory).
    packed: (w, w_lo) of pack_weights(conv1_1/filter as [27, 64], ksize 1, cin_pad 64); out: [N,H,W,64] bf16 or the
    hi/lo pair tensor [N,H,W,128]."""
    _chk_cuda(images, packed[0], packed[1], bias, out)
    N, H, W, _ = images.shape
    p = capi.Conv1Params(capi.ptr(images), N, H, W, capi.ptr(packed[0]), capi.ptr(packed[1]), capi.ptr(bias),
                         capi.ptr(out), capi.ptr(out, 64) if pair else None, out.shape[3], None, None, 0, None,
                         1 if pair else 0)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(capi.load().fcn8_conv1_fwd(C.byref(p), _stream()))
    if e0 is not None:
        TIMER.stop("conv1", 2.0 * N * H * W * 64 * 27, e0)
    return out


def conv1_wgrad(images, dy, dw, pair=False):
    """conv1_1/filter gradient dw [27, 64] (fp32, TF order) from the uint8 image and dy ([N,H,W,64] bf16 or pair)."""
    _chk_cuda(images, dy, dw)
    N, H, W, _ = images.shape
    p = capi.Conv1Params(capi.ptr(images), N, H, W, None, None, None, None, None, 0, capi.ptr(dy),
                         capi.ptr(dy, 64) if pair else None, dy.shape[3], capi.ptr(dw), 1 if pair else 0)
    lib = capi.load()
    nbytes = lib.fcn8_conv1_wgrad_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, images.device)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(lib.fcn8_conv1_wgrad(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    if e0 is not None:
        TIMER.stop("conv1", 2.0 * N * H * W * 64 * 27, e0)
    return dw


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


# ---------------------------------------------------------------------------------------------------- decoder
def deconv_cp(stride):
    """Channels per pixel of the padded blocked tensors of a stride-s transposed-convolution stage."""
    return capi.load().fcn8_deconv_cp(stride)


def head_pack(K, bias, split=True, out=None):
    """Score-head kernel K [Cin, C] (+ bias [C]) -> dict(w, w_lo: bf16 [Cin, 64] TF-layout weight with the classes
    zero-padded to 64; bias64: fp32 [64]).  `out`: an earlier result, refilled in place."""
    _chk_cuda(K, bias)
    cin, Cc = K.shape
    if out is None:
        out = dict(w=torch.empty((cin, 64), dtype=torch.bfloat16, device=K.device), w_lo=None,
                   bias64=torch.empty(64, dtype=torch.float32, device=K.device))
        if split:
            out["w_lo"] = torch.empty_like(out["w"])
    capi.check(capi.load().fcn8_head_pack(capi.ptr(K), capi.ptr(bias), cin, Cc, capi.ptr(out["w"]),
                                          capi.ptr(out["w_lo"]), capi.ptr(out["bias64"]), _stream()))
    return out


def deconv_pack(T, bias, stride, split=True, out=None):
    """T [2s,2s,C,C], bias [C] -> dict of phase-GEMM operands (w_fwd [s*s*CP, 256], w_dx [64, 4*s*s*CP] bf16 and their
    *_lo halves, bias_big fp32 [s*s*CP]).  `out`: a dict returned by an earlier call, refilled in place."""
    _chk_cuda(T, bias)
    Cc = T.shape[-1]
    ncols = stride * stride * deconv_cp(stride)
    if out is None:
        bf = dict(dtype=torch.bfloat16, device=T.device)
        out = dict(w_fwd=torch.empty((ncols, 256), **bf), w_dx=torch.empty((64, 4 * ncols), **bf),
                   bias_big=torch.empty(ncols, dtype=torch.float32, device=T.device), w_fwd_lo=None, w_dx_lo=None)
        if split:
            out["w_fwd_lo"] = torch.empty_like(out["w_fwd"])
            out["w_dx_lo"] = torch.empty_like(out["w_dx"])
    capi.check(capi.load().fcn8_deconv_pack(capi.ptr(T), capi.ptr(bias), Cc, stride, capi.ptr(out["w_fwd"]),
                                            capi.ptr(out["w_fwd_lo"]), capi.ptr(out["w_dx"]), capi.ptr(out["w_dx_lo"]),
                                            capi.ptr(out["bias_big"]), _stream()))
    return out


def planes_alloc(N, h, w, device, channels=64, zero=True):
    """A decoder activation as bf16 hi / lo planes: (hi, lo) tensors [N,h,w,channels] (zero beyond the class count)."""
    f = torch.zeros if zero else torch.empty
    return (f((N, h, w, channels), dtype=torch.bfloat16, device=device),
            f((N, h, w, channels), dtype=torch.bfloat16, device=device))


def planes_from_float(x, channels=64):
    """fp32 [N,h,w,C] -> (hi, lo) planes [N,h,w,channels] (test helper)."""
    N, h, w, Cc = x.shape
    hi, lo = planes_alloc(N, h, w, x.device, channels)
    hi[..., :Cc] = x.to(torch.bfloat16)
    lo[..., :Cc] = (x - hi[..., :Cc].float()).to(torch.bfloat16)
    return hi, lo


def planes_to_float(planes, num_classes):
    return planes[0][..., :num_classes].float() + planes[1][..., :num_classes].float()


def padded_alloc(N, h, w, stride, device):
    """Zero-bordered padded blocked planes [N, s*(h+1), s*(w+1), CP] of a stride-s stage (gradient of its output)."""
    cp = deconv_cp(stride)
    return planes_alloc(N, stride * (h + 1), stride * (w + 1), device, cp)


def padded_interior(planes, stride):
    """[N, s*h, s*w, CP] views of the interior of padded blocked planes."""
    p = stride // 2
    return tuple(t[:, p:t.shape[1] - p, p:t.shape[2] - p, :] for t in planes)


def _deconv_params(x=None, w=None, w_lo=None, bias_big=None, out=None, skip=None, dz=None, dT=None, colsum=None,
                   colsum_n=0, N=0, h=0, wd=0, Cc=0, stride=0, nseg=3, **loss):
    """x / out / skip / dz: (hi, lo) tuples of NHWC views."""
    x_hi, x_lo = x if x is not None else (None, None)
    o_hi, o_lo = out if out is not None else (None, None)
    s_hi, s_lo = skip if skip is not None else (None, None)
    z_hi, z_lo = dz if dz is not None else (None, None)
    _chk_view(x_hi, x_lo, o_hi, o_lo, s_hi, s_lo)
    _chk_cuda(z_hi, z_lo, w, w_lo)
    x_ld, x_sH, x_sN = _geom(x_hi) if x_hi is not None else (0, 0, 0)
    o_ld, o_sH, o_sN = _geom(o_hi) if o_hi is not None else (0, 0, 0)
    if s_hi is not None and _geom(s_hi) != _geom(o_hi):
        raise capi.Fcn8Error("skip planes must have the geometry of out")
    dzo = loss.get("dz_out") or (None, None)
    return capi.DeconvParams(capi.ptr(x_hi), capi.ptr(x_lo), x_ld, x_sH, x_sN, capi.ptr(w), capi.ptr(w_lo),
                             capi.ptr(bias_big), capi.ptr(o_hi), capi.ptr(o_lo), o_ld, o_sH, o_sN, capi.ptr(s_hi),
                             capi.ptr(s_lo), capi.ptr(z_hi), capi.ptr(z_lo), capi.ptr(dT), capi.ptr(colsum), colsum_n,
                             N, h, wd, Cc, stride, nseg, capi.ptr(loss.get("labels")), capi.ptr(loss.get("loss_sum")),
                             capi.ptr(loss.get("dbias")), capi.ptr(dzo[0]), capi.ptr(dzo[1]),
                             capi.ptr(loss.get("logits")), capi.ptr(loss.get("softmax")), capi.ptr(loss.get("argmax")),
                             capi.ptr(loss.get("conf")), loss.get("grad_scale", 1.0), capi.ptr(loss.get("argmax_u8")))


def _deconv_flops(N, h, w, stride, Cc):
    return 2.0 * N * h * w * 4 * stride * stride * Cc * Cc


def deconv_fwd(x, packed, num_classes, stride, out, skip=None, nseg=3):
    """Stride-2 transposed convolution of planes x [N,h,w,.] into dense planes out [N,2h,2w,.] (+ skip planes)."""
    N, h, w, _ = x[0].shape
    p = _deconv_params(x=x if nseg == 3 else (x[0], None), w=packed["w_fwd"], w_lo=packed["w_fwd_lo"] if nseg > 1 else None,
                       bias_big=packed["bias_big"], out=out, skip=skip, N=N, h=h, wd=w, Cc=num_classes, stride=stride,
                       nseg=nseg)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(capi.load().fcn8_deconv_fwd(C.byref(p), _stream()))
    if e0 is not None:
        TIMER.stop("deconv", _deconv_flops(N, h, w, stride, num_classes), e0)
    return out


def deconv_loss(x, packed, num_classes, nseg=3, labels=None, loss_sum=None, dz_out=None, dbias=None, grad_scale=1.0,
                logits=None, softmax=None, argmax=None, conf=None, argmax_u8=None):
    """upscore8 (stride 8) with the loss / predictor fused into its epilogue; see fcn8_deconv_loss."""
    N, h, w, _ = x[0].shape
    _chk_cuda(labels, loss_sum, dbias, logits, softmax, argmax, conf, argmax_u8)
    if dz_out is not None:
        _chk_cuda(dz_out[0], dz_out[1])
    p = _deconv_params(x=x if nseg == 3 else (x[0], None), w=packed["w_fwd"], w_lo=packed["w_fwd_lo"] if nseg > 1 else None,
                       bias_big=packed["bias_big"], N=N, h=h, wd=w, Cc=num_classes, stride=8, nseg=nseg, labels=labels,
                       loss_sum=loss_sum, dz_out=dz_out, dbias=dbias, grad_scale=grad_scale, logits=logits,
                       softmax=softmax, argmax=argmax, conf=conf, argmax_u8=argmax_u8)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(capi.load().fcn8_deconv_loss(C.byref(p), _stream()))
    if e0 is not None:
        px = N * 64 * h * w   # HBM-side work: algorithmic bytes of the requested outputs + the input planes
        nbytes = N * h * w * num_classes * 4 + (px * num_classes if labels is not None else 0) + \
            (px * num_classes * 4 if dz_out is not None else 0) + (px * num_classes * 4 if logits is not None else 0) + \
            (px * num_classes * 4 if softmax is not None else 0) + (px * 8 if argmax is not None else 0) + \
            (px if argmax_u8 is not None else 0)
        TIMER.stop("upscore8_fused", float(nbytes), e0)


def deconv_dx(dz, packed, num_classes, stride, out, nseg=3, colsum=None, shape=None):
    """Input gradient of a stride-s stage from padded blocked dz planes into planes `out` [N,h,w,.] (may be the
    interior view of the next stage's padded planes); colsum[c] += sum over pixels of the result."""
    N, h, w, _ = out[0].shape
    p = _deconv_params(dz=dz if nseg == 3 else (dz[0], None), w=packed["w_dx"], w_lo=packed["w_dx_lo"] if nseg > 1 else None,
                       out=out, colsum=colsum, N=N, h=h, wd=w, Cc=num_classes, stride=stride, nseg=nseg)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(capi.load().fcn8_deconv_dx(C.byref(p), _stream()))
    if e0 is not None:
        TIMER.stop("deconv", _deconv_flops(N, h, w, stride, num_classes), e0)
    return out


def deconv_dw(x, dz, num_classes, stride, dT, nseg=3):
    """Filter gradient dT [2s,2s,C,C] (TF layout, fp32) of a stride-s stage from planes x and padded blocked dz planes."""
    N, h, w, _ = x[0].shape
    _chk_cuda(dT)
    p = _deconv_params(x=x if nseg == 3 else (x[0], None), dz=dz if nseg > 1 else (dz[0], None), dT=dT, N=N, h=h, wd=w,
                       Cc=num_classes, stride=stride, nseg=nseg)
    lib = capi.load()
    nbytes = lib.fcn8_deconv_dw_workspace_bytes(C.byref(p))
    ws = _workspace(nbytes, x[0].device)
    e0 = TIMER.start() if TIMER is not None else None
    capi.check(lib.fcn8_deconv_dw(C.byref(p), capi.ptr(ws), nbytes, _stream()))
    if e0 is not None:
        TIMER.stop("deconv", _deconv_flops(N, h, w, stride, num_classes), e0)
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


def pack_labels(onehot, ids, threads=4):
    """HOST code: one-hot uint8 / bool numpy batch [..., C] -> class ids (uint8 numpy / pinned buffer `ids`, one per
    pixel).  Returns True when every pixel was exactly one-hot (ids is valid), False otherwise."""
    import numpy as np
    a = onehot.view(np.uint8) if onehot.dtype == np.bool_ else onehot
    if a.dtype != np.uint8 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("pack_labels: contiguous uint8 / bool labels expected")
    Cn = a.shape[-1]
    pixels = a.size // Cn
    if ids.size < pixels or ids.dtype != np.uint8 or not ids.flags["C_CONTIGUOUS"]:
        raise ValueError("pack_labels: ids must be a contiguous uint8 buffer of one byte per pixel")
    rc = capi.load().fcn8_pack_labels(a.ctypes.data, pixels, Cn, ids.ctypes.data, threads)
    if rc < 0:
        capi.check(rc)
    return rc == 0


def expand_labels(ids, onehot):
    """onehot (uint8 CUDA tensor [..., C]) = one-hot of the class ids (uint8 CUDA tensor, one per pixel)."""
    _chk_cuda(ids, onehot)
    Cn = onehot.shape[-1]
    capi.check(capi.load().fcn8_expand_labels(capi.ptr(ids), capi.ptr(onehot), onehot.numel() // Cn, Cn, _stream()))
    return onehot


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
