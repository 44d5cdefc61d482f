"""FCN-8s forward / backward / Adam engine on libfcn8s_sm100.so.

This is the replacement for what `tf.Session.run` executes in the reference: the external VGG-16 encoder graph
(fcn8s_tensorflow.py:127-152), `_build_decoder` (:154-237), `_build_optimizer` (:239-259) and `_build_predictor`
(:261-271).  torch owns the device memory (flat parameter / gradient / Adam buffers, a per-shape activation arena) and
the stream; every arithmetic operation is a hand-written sm_100a kernel reached through the C ABI
(include/fcn8s_b200.h).  There is no CPU path: constructing an Engine without a CUDA device or without the built
library raises.

Precision modes: see `Engine`.  The decoder (score heads, transposed convs, loss) is fp32 in every mode.
"""
import functools
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _capi as capi
from . import ops

VGG_BLOCKS = [(1, 64, 2), (2, 128, 2), (3, 256, 3), (4, 512, 3), (5, 512, 3)]
POOL3_SCALE, POOL4_SCALE = 1e-4, 1e-2          # fcn8s_tensorflow.py:171,182
BETA1, BETA2, EPS = 0.9, 0.999, 1e-8           # tf.train.AdamOptimizer defaults (:256)
# (packed-operand key, TF variable scope, stride) of upscore2 / upscore_pool4 / upscore8 (fcn8s_tensorflow.py:204-233)
UPSCORE_STAGES = [("up2", "fc7_conv2d_trans", 2), ("up4", "fc7_pool4_conv2d_trans", 2),
                  ("up8", "fc7_pool4_pool3_conv2d_trans", 8)]
DECODER_KERNELS = ["pool3_1x1/kernel", "pool4_1x1/kernel", "fc7_1x1/kernel", "fc7_conv2d_trans/kernel",
                   "fc7_pool4_conv2d_trans/kernel", "fc7_pool4_pool3_conv2d_trans/kernel"]


def encoder_layers():
    """[(name, ksize, cin, cout)] in forward order, fc6/fc7 included."""
    out = []
    cin = 3
    for b, cout, n in VGG_BLOCKS:
        for i in range(1, n + 1):
            out.append(("conv%d_%d" % (b, i), 3, cin, cout))
            cin = cout
    out.append(("fc6", 7, 512, 4096))
    out.append(("fc7", 1, 4096, 4096))
    return out


def _wname(layer):
    return layer + ("/weights" if layer.startswith("fc") else "/filter")


def variable_shapes(num_classes):
    """name -> TF-layout shape of the 42 trainable variables, forward order."""
    C = num_classes
    s = OrderedDict()
    for name, k, cin, cout in encoder_layers():
        s[_wname(name)] = (k, k, cin, cout)
        s[name + "/biases"] = (cout,)
    s["pool3_1x1/kernel"] = (1, 1, 256, C)
    s["pool3_1x1/bias"] = (C,)
    s["pool4_1x1/kernel"] = (1, 1, 512, C)
    s["pool4_1x1/bias"] = (C,)
    s["fc7_1x1/kernel"] = (1, 1, 4096, C)
    s["fc7_1x1/bias"] = (C,)
    s["fc7_conv2d_trans/kernel"] = (4, 4, C, C)
    s["fc7_conv2d_trans/bias"] = (C,)
    s["fc7_pool4_conv2d_trans/kernel"] = (4, 4, C, C)
    s["fc7_pool4_conv2d_trans/bias"] = (C,)
    s["fc7_pool4_pool3_conv2d_trans/kernel"] = (16, 16, C, C)
    s["fc7_pool4_pool3_conv2d_trans/bias"] = (C,)
    return s


def flat_layout(num_classes):
    """Flat-buffer layout in backward-completion order (decoder | fc7 W | fc6 W | conv5_3 W ... conv1_1 W | encoder
    biases fc7 ... conv1_1) so that a gradient all-reduce can start on the big fc6/fc7 blocks while the conv backward
    still runs.  The encoder biases form one contiguous block at the end: their gradients are accumulated with atomics
    by the fused dgrad / max-pool-backward epilogues, so the block is zeroed with one fill per step.  Every tensor
    starts on a 256-byte boundary.  Returns (OrderedDict name -> (offset_elems, shape), total_elems)."""
    shapes = variable_shapes(num_classes)
    order = []
    for base in ["fc7_pool4_pool3_conv2d_trans", "pool3_1x1", "fc7_pool4_conv2d_trans", "pool4_1x1",
                 "fc7_conv2d_trans", "fc7_1x1"]:
        order += [base + "/kernel", base + "/bias"]
    for name, _, _, _ in reversed(encoder_layers()):
        order.append(_wname(name))
    for name, _, _, _ in reversed(encoder_layers()):
        order.append(name + "/biases")
    assert set(order) == set(shapes)
    layout = OrderedDict()
    off = 0
    for n in order:
        layout[n] = (off, shapes[n])
        off += int(np.prod(shapes[n]))
        off = (off + 63) // 64 * 64
    return layout, off


def _on_device(fn):
    """Run an Engine method with the engine's device current: every launch goes to `torch.cuda.current_stream()` and
    the library's runtime calls target the current device, so an Engine built with device='cuda:1' must not depend on
    the caller having called `torch.cuda.set_device(1)`."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)
    return wrapper


class Engine:
    """Precision modes (`precision=`), all with fp32 accumulation in TMEM, fp32 master weights / gradients / Adam:

      "bf16"    activations + tensor-core operands bf16.
      "fp32"    fp32-equivalent: activations and weights are bf16 hi/lo PAIRS (FCN8_BF16X2: v ~ hi + lo, 16-17
                mantissa bits) and every GEMM forms the error-compensated hi*hi + hi*lo + lo*hi on the bf16 tensor
                cores (3 MMAs per product at the bf16 rate = half the cost of 3xTF32).  Meets the 1e-4 logit tolerance.
      "tf32x3"  fp32 activations, 3xTF32 error-compensated products (kind::tf32, 6x the bf16 cost).
      "tf32"    fp32 activations, single tf32 pass.
    In "bf16" / "fp32" the GEMMs read a bf16 (hi/lo) shadow of the flat parameter buffer in its TF layout directly
    (refreshed by the Adam kernel), so there is no per-step weight packing."""

    MODES = ("bf16", "fp32", "tf32x3", "tf32")

    def __init__(self, num_classes, precision="bf16", device=None, backward_terms=3, grad_comm=None):
        """backward_terms (precision "fp32" only): bf16 products per algorithmic product in the BACKWARD GEMMs (dgrad and
        filter gradients of the encoder): 3 = hi*hi + hi*lo + lo*hi like the forward pass (default); 2 / 1 are the
        measured, non-default "fp32 forward / reduced backward" modes (one operand, or both, rounded to bf16).
        grad_comm: wire format of the data-parallel gradient all-reduce, "fp32" or "bf16" (default: bf16 in the bf16
        precision mode, fp32 otherwise)."""
        if not torch.cuda.is_available():
            raise capi.Fcn8Error("fcn8s_tensorflow_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if precision not in self.MODES:
            raise ValueError("precision must be one of %s" % (self.MODES,))
        if not (1 <= num_classes <= 32):
            raise ValueError("num_classes must be in [1, 32]")
        self.lib = capi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        capi.check(self.lib.fcn8_device_check(self.device.index))
        if backward_terms not in (1, 2, 3):
            raise ValueError("backward_terms must be 1, 2 or 3")
        self.bt = int(backward_terms) if precision == "fp32" else 3
        self.grad_comm = grad_comm or os.environ.get("FCN8_GRAD_COMM") or ("bf16" if precision == "bf16" else "fp32")
        if self.grad_comm not in ("fp32", "bf16"):
            raise ValueError("grad_comm must be 'fp32' or 'bf16'")
        self.rank = 0
        self.g16 = None          # bf16 wire copy of the flat gradient (allocated by dist.attach when grad_comm == bf16)
        self.C = num_classes
        self.precision = precision
        self.pair = precision == "fp32"                      # bf16 hi/lo pair activations
        self.hwio = precision in ("bf16", "fp32")            # GEMMs read the TF-layout weight shadow
        self.x3 = precision == "tf32x3"                      # legacy 3xTF32 encoder products
        self.dec3 = precision in ("fp32", "tf32x3")          # 3xTF32 in the (fp32) decoder GEMMs
        self.dt = ops.BF16 if self.hwio else ops.F32         # tensor-core operand type of the encoder
        self.tdt = torch.bfloat16 if self.hwio else torch.float32
        self.cm = 2 if self.pair else 1                      # stored channels per logical channel
        self.kp = 64 if self.hwio else 32                    # padded im2col width of conv1_1
        self.rnd = ops.EPI_ROUND_TF32 if precision == "tf32" else 0
        self.layers = encoder_layers()
        self.layout, self.n_flat = flat_layout(num_classes)
        self.bias_block = self.layout["fc7/biases"][0]       # encoder biases: [bias_block, n_flat)
        z = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(self.n_flat, **z)
        self.grads = torch.zeros(self.n_flat, **z)
        self.adam_m = torch.zeros(self.n_flat, **z)
        self.adam_v = torch.zeros(self.n_flat, **z)
        self.w_hi = torch.zeros(self.n_flat, dtype=torch.bfloat16, device=self.device) if self.hwio else None
        self.w_lo = torch.zeros(self.n_flat, dtype=torch.bfloat16, device=self.device) if self.pair else None
        self.global_step = 0
        self.loss_buf = torch.zeros(2, **z)   # [0] = sum of per-pixel CE, [1] = L2 regularisation loss
        self.packed = {}
        self._packed_dirty = True
        self._shadow_dirty = True
        self._arenas = {}
        self.world = 1
        self.allreduce = None  # dist.GradientAllReduce installed by the data-parallel wrapper
        self._reduced_upto = 0  # elements of self.grads whose all-reduce has already been started this step
        # per-step scalars in device memory ([0] lr_t, [1] dropout seed bits): a captured step can be replayed
        self.step_scalars = torch.zeros(4, **z)
        self.use_graphs = os.environ.get("FCN8_GRAPHS", "1") != "0"
        self._graphs = {}           # (shape, keep_prob, l2_rate) -> (CUDAGraph, kernel launches per replay)
        self._warm = {}
        self.graph_launches = 0     # kernels launched through graph replays (fcn8_launch_count() does not see those)
        self.head_elems = self.layout["conv5_3/filter"][0]   # [decoder | fc7 W | fc6 W] prefix of the flat buffer
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.dp_reserve_sms = int(os.environ.get("FCN8_DP_RESERVE_SMS", "0"))

    # ------------------------------------------------------------------ parameters
    def view(self, name, buf=None):
        off, shape = self.layout[name]
        buf = self.params if buf is None else buf
        return buf[off:off + int(np.prod(shape))].view(shape)

    @_on_device
    def load_weights(self, weights):
        """weights: mapping TF variable name -> array/tensor in TF layout. Missing names raise KeyError."""
        for name in self.layout:
            w = weights[name]
            t = torch.as_tensor(np.asarray(w) if not torch.is_tensor(w) else w).to(torch.float32)
            if tuple(t.shape) != tuple(self.layout[name][1]):
                raise ValueError("shape mismatch for %s: %s vs %s" % (name, tuple(t.shape), self.layout[name][1]))
            self.view(name).copy_(t.to(self.device))
        self._packed_dirty = True
        self._shadow_dirty = True

    @_on_device
    def state_dict(self):
        return OrderedDict((n, self.view(n).detach().cpu().clone()) for n in self.layout)

    @_on_device
    def grad_dict(self):
        return OrderedDict((n, self.view(n, self.grads).detach().cpu().clone()) for n in self.layout)

    @_on_device
    def repack(self):
        """Derived tensor-core operands of the parameters.  hwio modes: only conv1_1 (27 -> 64 im2col columns) and the
        upscore8 phase-GEMM operands are packed (the rest is the shadow written by Adam); legacy tf32 modes: fprop and
        dgrad operand copies of every encoder layer."""
        if self.hwio and self._shadow_dirty:
            ops.shadow_weights(self.params, self.w_hi, self.w_lo)
            self._shadow_dirty = False
        for name, k, cin, cout in self.layers:
            w = self.view(_wname(name))
            old = self.packed.get(name)   # refilled in place: a captured CUDA graph keeps reading the same buffers
            if name == "conv1_1":
                # runs as a 1x1 conv over the 27-column im2col (padded to kp) built by the feed kernel
                self.packed[name] = ops.pack_weights(w, 1, 27, cout, 0, self.dt, cin_pad=self.kp,
                                                     split=self.x3 or self.pair,
                                                     out=old[0:2] if old else None) + (None, None)
            elif not self.hwio:
                f = ops.pack_weights(w, k, cin, cout, 0, self.dt, split=self.x3, out=old[0:2] if old else None)
                d = ops.pack_weights(w, k, cin, cout, 1, self.dt, split=self.x3, out=old[2:4] if old else None)
                self.packed[name] = f + d
        for key, base, stride in UPSCORE_STAGES:   # phase-GEMM operands of the three transposed convolutions
            self.packed[key] = ops.upscore_tc_pack(self.view(base + "/kernel"), self.view(base + "/bias"), stride,
                                                   split=self.dec3, out=self.packed.get(key))
        self._packed_dirty = False

    # ------------------------------------------------------------------ activation arena
    def _arena(self, N, H, W):
        key = (N, H, W)
        a = self._arenas.get(key)
        if a is None:
            if H % 32 or W % 32:
                raise ValueError("image height and width must be multiples of 32 (got %dx%d): the reference graph's "
                                 "skip additions (fcn8s_tensorflow.py:213,224) only align for such sizes" % (H, W))
            a = {}
            self._arenas[key] = a
        return a

    def _buf(self, arena, name, shape, dtype):
        t = arena.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            arena[name] = t
        return t

    def _act(self, arena, name, N, h, w, c):
        """Encoder activation buffer [N,h,w,c] in the mode's storage format ([N,h,w,2c] bf16 for hi/lo pairs)."""
        return self._buf(arena, name, (N, h, w, c * self.cm), self.tdt)

    def _split(self, arena, name, x, force=False):
        """3xTF32 operands of an fp32 tensor x: (x, lo) with lo = round_tf32(x - trunc_tf32(x)) in an arena buffer --
        the tensor core truncates x to its tf32 high part by itself; (x, None) when the mode is single-pass."""
        if not (self.x3 or force):
            return x, None
        lo = self._buf(arena, name + ".lo", x.shape, torch.float32)
        capi.check(self.lib.fcn8_split_tf32(capi.ptr(x), None, capi.ptr(lo), x.numel(), ops._stream()))
        return x, lo

    def _wview(self, name):
        """(hi, lo) bf16 shadow views of an encoder weight tensor in TF layout."""
        return self.view(_wname(name), self.w_hi), (self.view(_wname(name), self.w_lo) if self.pair else None)

    def _conv(self, A, name, x, k, cin, cout, out, bias, flags, **kw):
        """Encoder convolution of layer `name` (forward direction) in the mode's operand format."""
        if name == "conv1_1":
            wp, wlo = self.packed[name][0], self.packed[name][1]
            xl = self._split(A, "in_" + name, x)[1] if self.x3 else None
            return ops.conv_gemm(x, wp, cout, 1, bias=bias, flags=flags | self.rnd, out=out, x_lo=xl, wp_lo=wlo,
                                 pair=self.pair, **kw)
        if self.hwio:
            wh, wl = self._wview(name)
            return ops.conv_gemm(x, wh, cout, k, bias=bias, flags=flags, out=out, wp_lo=wl, pair=self.pair, w_mode=1,
                                 **kw)
        wp, wlo = self.packed[name][0], self.packed[name][1]
        xl = self._split(A, "in_" + name, x)[1]
        return ops.conv_gemm(x, wp, cout, k, bias=bias, flags=flags | self.rnd, out=out, x_lo=xl, wp_lo=wlo, **kw)

    # ------------------------------------------------------------------ forward
    @_on_device
    def forward(self, images, keep_prob=1.0, seed=0, train=False, _scalars_set=False):
        """images: uint8 CUDA tensor [N,H,W,3] (RGB). Returns fp32 logits [N,H,W,C] (a view of an arena buffer).
        The dropout seed is read by the kernels from `step_scalars` (set here unless the caller already did)."""
        if self._packed_dirty or (self.hwio and self._shadow_dirty):
            self.repack()
        if train and keep_prob < 1.0 and not _scalars_set:
            ops.set_step_scalars(self.step_scalars, 0.0, seed)
        N, H, W, _ = images.shape
        A = self._arena(N, H, W)
        A["_shape"] = (N, H, W)
        pair = self.pair
        x = self._act(A, "im2col", N, H, W, self.kp)
        p = capi.PreprocessParams(capi.ptr(images), capi.ptr(x), N, H, W, ops.BF16X2 if pair else self.dt)
        capi.check(self.lib.fcn8_preprocess_im2col(ops.C.byref(p), ops._stream()))
        if self.precision == "tf32":   # single-pass tf32: operands pre-rounded (the MMA truncates)
            capi.check(self.lib.fcn8_split_tf32(capi.ptr(x), capi.ptr(x), None, x.numel(), ops._stream()))
        h, w = H, W
        li = 0
        for b, cout, n in VGG_BLOCKS:
            for i in range(1, n + 1):
                name, k, cin, _ = self.layers[li]
                li += 1
                out = self._act(A, name, N, h, w, cout)
                self._conv(A, name, x, k, cin, cout, out, self.view(name + "/biases"), ops.EPI_BIAS | ops.EPI_RELU)
                x = out
            h, w = (h + 1) // 2, (w + 1) // 2
            pooled = self._act(A, "pool%d" % b, N, h, w, cout)
            ops.maxpool_fwd(x, out=pooled, pair=pair)
            x = pooled
        drop = train and keep_prob < 1.0
        for name, k, cin, cout in self.layers[-2:]:
            out = self._act(A, name, N, h, w, cout)
            flags = ops.EPI_BIAS | ops.EPI_RELU | (ops.EPI_DROPOUT if drop else 0)
            self._conv(A, name, x, k, cin, cout, out, self.view(name + "/biases"), flags,
                       keep_prob=keep_prob if drop else 1.0, seed=1 if name == "fc7" else 0,
                       seed_ptr=self.step_scalars[1:2] if drop else None)
            x = out
        C = self.C
        f32 = torch.float32
        s3 = ops.score_head_fwd(A["pool3"], self.view("pool3_1x1/kernel").view(256, C), self.view("pool3_1x1/bias"),
                                POOL3_SCALE, out=self._buf(A, "s3", (N, H // 8, W // 8, C), f32), pair=pair)
        s4 = ops.score_head_fwd(A["pool4"], self.view("pool4_1x1/kernel").view(512, C), self.view("pool4_1x1/bias"),
                                POOL4_SCALE, out=self._buf(A, "s4", (N, H // 16, W // 16, C), f32), pair=pair)
        s7 = ops.score_head_fwd(A["fc7"], self.view("fc7_1x1/kernel").view(4096, C), self.view("fc7_1x1/bias"), 1.0,
                                out=self._buf(A, "s7", (N, H // 32, W // 32, C), f32), pair=pair)
        # the three transposed convolutions as tcgen05 phase GEMMs over padded blocked tensors; the skip adds
        # (fcn8s_tensorflow.py:213,224) ride on the gather of the block interior
        ld4 = (C + 3) // 4 * 4
        f4 = self._buf(A, "f4", (N, H // 16, W // 16, ld4), f32)
        self._upscore_stage_fwd(A, "up2", self._pad4(A, "s7p", s7), 2, skip=s4, out=f4)
        f3 = self._buf(A, "f3", (N, H // 8, W // 8, ld4), f32)
        self._upscore_stage_fwd(A, "up4", f4, 2, skip=s3, out=f3)
        zp = self._upscore_stage_fwd(A, "up8", f3, 8)
        A["logits_p"] = zp
        A["logits"] = ops.upscore_tc_interior(zp, C, 8)
        return A["logits"]

    def _upscore_stage_fwd(self, A, key, x, stride, skip=None, out=None):
        """One transposed convolution: x [N,h,w,ld4] -> padded blocked A[key + "_p"]; with `out`, also the interior
        (+ skip) as the next stage's dense input."""
        N, h, w, _ = x.shape
        zp = A.get(key + "_p")
        if zp is None:
            zp = A[key + "_p"] = ops.upscore_tc_alloc(N, h, w, self.C, stride, self.device)
        x_lo = self._split(A, key + "_x", x, force=self.dec3)[1]
        ops.upscore_tc_fwd(x, self.packed[key], self.C, stride, zp, x_lo=x_lo)
        if out is not None:
            ops.upscore_tc_gather(zp, skip, out, self.C, stride)
        return zp

    def _upscore_stage_bwd(self, A, key, base, x, g, stride, dzp=None):
        """Backward of one transposed convolution: g = gradient of its (dense) output, or dzp already in the padded
        blocked layout (upscore8: written by the loss kernel).  Fills dT / dbias in the flat gradient, returns dx."""
        N, h, w, ld = x.shape
        G = self.grads
        if dzp is None:
            dzp = A.get(key + "_dp")
            if dzp is None:   # zero border / pad channels, never written again
                dzp = A[key + "_dp"] = ops.upscore_tc_alloc(N, h, w, self.C, stride, self.device, zero=True)
            db = self.view(base + "/bias", G)
            db.zero_()
            ops.upscore_tc_scatter(g, dzp, self.C, stride, dbias=db)
        dz_lo = self._split(A, key + "_dp", dzp, force=self.dec3)[1]
        ops.upscore_tc_dw(x, dzp, self.C, stride, self.view(base + "/kernel", G), x_lo=A.get(key + "_x.lo"),
                          dzp_lo=dz_lo)
        dx = self._buf(A, key + "_dx", x.shape, torch.float32)
        ops.upscore_tc_dx(dzp, self.packed[key], self.C, stride, dx, dzp_lo=dz_lo)
        return dx

    def _dense(self, A, name, t):
        """[..., ld4] -> dense [..., C] copy for the score-head kernels (a no-op view when C is a multiple of 4)."""
        if t.shape[-1] == self.C:
            return t
        d = self._buf(A, name, tuple(t.shape[:-1]) + (self.C,), torch.float32)
        d.copy_(t[..., :self.C])
        return d

    def _pad4(self, arena, name, t):
        """[N,h,w,C] -> the same tensor with the channel stride rounded up to a multiple of 4 (zero filled): the layout
        the TMA maps of the decoder GEMMs need. No-op (no copy) when C is already a multiple of 4."""
        Cc = t.shape[-1]
        if Cc % 4 == 0:
            return t
        buf = arena.get(name)
        if buf is None:
            buf = arena[name] = torch.zeros(tuple(t.shape[:-1]) + ((Cc + 3) // 4 * 4,), dtype=t.dtype, device=t.device)
        buf[..., :Cc].copy_(t)
        return buf

    @staticmethod
    def dropout_seed(seed, layer):
        """Seed of the counter-based dropout mask of `layer` for step seed `seed` (host replica: rng.py); the kernels
        form the same value from the device scalar: *seed_ptr * 2 + (layer == fc7)."""
        return (int(seed) * 2 + (1 if layer == "fc7" else 0)) & 0xFFFFFFFF

    # ------------------------------------------------------------------ loss + backward
    @_on_device
    def loss_and_backward(self, images, labels, keep_prob=1.0, l2_rate=0.0, seed=0, _scalars_set=False):
        """Forward + backward of optimizer/total_loss (fcn8s_tensorflow.py:250-257) into self.grads.
        labels: uint8/bool CUDA tensor [N,H,W,C] one-hot. Returns the device scalar pair loss_buf (CE sum, L2)."""
        N, H, W, _ = images.shape
        C = self.C
        pair = self.pair
        self.forward(images, keep_prob, seed, train=True, _scalars_set=_scalars_set)
        A = self._arena(N, H, W)
        G = self.grads
        f32 = torch.float32
        self.loss_buf.zero_()
        G[self.bias_block:].zero_()   # encoder bias gradients are accumulated by fused epilogues (atomics)
        zp = A["logits_p"]
        dzp = A.get("dlogits_p")
        if dzp is None:   # zero border, never written again
            dzp = A["dlogits_p"] = ops.upscore_tc_alloc(N, H // 8, W // 8, C, 8, self.device, zero=True)
        npx = N * H * W
        dc8 = self.view("fc7_pool4_pool3_conv2d_trans/bias", G)
        dc8.zero_()
        ops.softmax_xent(zp, labels.view(torch.uint8), self.loss_buf[0:1], dzp, grad_scale=1.0 / npx, dbias=dc8,
                         pad=4, num_classes=C)
        # decoder backward: three phase-GEMM stages; the skip adds fan the gradient out
        df3p = self._upscore_stage_bwd(A, "up8", "fc7_pool4_pool3_conv2d_trans", A["f3"], None, 8, dzp=dzp)
        df4p = self._upscore_stage_bwd(A, "up4", "fc7_pool4_conv2d_trans", A["f4"], df3p, 2)
        ds7p = self._upscore_stage_bwd(A, "up2", "fc7_conv2d_trans", A["s7p"] if C % 4 else A["s7"], df4p, 2)
        df3, df4, ds7 = self._dense(A, "df3", df3p), self._dense(A, "df4", df4p), self._dense(A, "ds7", ds7p)
        inv_keep = 1.0 / keep_prob if keep_prob < 1.0 else 1.0
        # score heads: ds3 = df3, ds4 = df4 (the adds fan the gradient out unchanged)
        dpool3 = self._buf(A, "d_pool3_head", A["pool3"].shape, self.tdt)
        ops.score_head_bwd(A["pool3"], self.view("pool3_1x1/kernel").view(256, C), df3, POOL3_SCALE,
                           self.view("pool3_1x1/kernel", G).view(256, C), self.view("pool3_1x1/bias", G), dpool3,
                           pair=pair)
        dpool4 = self._buf(A, "d_pool4_head", A["pool4"].shape, self.tdt)
        ops.score_head_bwd(A["pool4"], self.view("pool4_1x1/kernel").view(512, C), df4, POOL4_SCALE,
                           self.view("pool4_1x1/kernel", G).view(512, C), self.view("pool4_1x1/bias", G), dpool4,
                           pair=pair)
        # fc7 output: dropout + ReLU backward folded into the head's dx (mask = fc7 > 0, scale 1/keep_prob)
        dy = self._buf(A, "d_fc7", A["fc7"].shape, self.tdt)
        ops.score_head_bwd(A["fc7"], self.view("fc7_1x1/kernel").view(4096, C), ds7, 1.0,
                           self.view("fc7_1x1/kernel", G).view(4096, C), self.view("fc7_1x1/bias", G), dy,
                           mask=True, mask_scale=inv_keep, pair=pair)
        if l2_rate != 0.0:
            for kname in DECODER_KERNELS:
                ops.l2_reg(self.view(kname).reshape(-1), self.view(kname, G).reshape(-1), self.loss_buf[1:2], l2_rate)
        # encoder backward
        for li in range(len(self.layers) - 1, -1, -1):
            name, k, cin, cout = self.layers[li]
            x_in = self._layer_input(A, li)
            gw = self.view(_wname(name), G)
            if name == "fc7":   # every other bias gradient is fused into the kernel that produces that layer's dY
                ops.bias_grad(dy, self.view(name + "/biases", G), pair=pair)
            dyh, dyl = self._split(A, "dy_" + name, dy)
            xl = A["in_%s.lo" % name] if self.x3 else None
            if name == "conv1_1":
                ops.wgrad_gemm(x_in, dyh, 1, gw.view(27, cout), rows_valid=27, x_lo=xl, dy_lo=dyl, pair=pair,
                               nseg=self.bt)
                break
            ops.wgrad_gemm(x_in, dyh, k, gw.view(k * k * cin, cout), x_lo=xl, dy_lo=dyl, pair=pair, nseg=self.bt)
            if name == "fc6" and self.allreduce is not None and getattr(self.allreduce, "overlap", False):
                # decoder, fc7 and fc6 gradients (89 % of the buffer) are final: reduce them under the conv backward
                self._start_reduce(0, self.head_elems)
                self._reduced_upto = self.head_elems
                if self.dp_reserve_sms > 0:   # leave SMs to the collective's CTAs while it runs under the backward
                    self.lib.fcn8_set_sm_limit(self.sm_count - self.dp_reserve_sms)
            dx = self._buf(A, "dx_" + name, x_in.shape, self.tdt)
            prev_name = self.layers[li - 1][0]
            prev_db = self.view(prev_name + "/biases", G)
            if self.hwio:
                wh, wl = self._wview(name)
                wkw = dict(wp_lo=wl, pair=pair, w_mode=2, nseg=self.bt)
            else:
                wh = self.packed[name][2]
                wkw = dict(x_lo=dyl, wp_lo=self.packed[name][3])
            if self._input_is_pool(li):
                # x_in is a pool output: no ReLU mask here (the pool backward applies it); add the score-head
                # gradient at pool3 / pool4 (AddN of the two consumers)
                pool_idx = self._pool_index(li)
                res = dpool3 if pool_idx == 3 else (dpool4 if pool_idx == 4 else None)
                ops.conv_gemm(dyh, wh, cin, k, flags=(ops.EPI_RESIDUAL if res is not None else 0) | self.rnd,
                              residual=res, out=dx, **wkw)
                src = A[prev_name]  # pre-pool activation (post-ReLU)
                dpre = self._buf(A, "dpre_" + prev_name, src.shape, self.tdt)
                ops.maxpool_bwd(src, dx, out=dpre, pair=pair, db=prev_db)
                dy = dpre
            else:
                # ReLU (and for fc6 -> dropout) backward of the producer fused as an epilogue mask on its output
                scale = inv_keep if prev_name == "fc6" else 1.0
                ops.conv_gemm(dyh, wh, cin, k, flags=ops.EPI_MASK | self.rnd, mask_src=x_in, mask_scale=scale, out=dx,
                              colsum=prev_db, **wkw)
                dy = dx
        if self.dp_reserve_sms > 0:
            self.lib.fcn8_set_sm_limit(0)
        return self.loss_buf

    def _input_is_pool(self, li):
        name = self.layers[li][0]
        return name == "fc6" or (name.startswith("conv") and name.endswith("_1") and name != "conv1_1")

    def _pool_index(self, li):
        name = self.layers[li][0]
        return 5 if name == "fc6" else int(name[4]) - 1

    def _layer_input(self, A, li):
        name = self.layers[li][0]
        if name == "conv1_1":
            return A["im2col"]
        if self._input_is_pool(li):
            return A["pool%d" % self._pool_index(li)]
        return A[self.layers[li - 1][0]]

    # ------------------------------------------------------------------ optimiser
    def _lr_t(self, lr, t):
        return float(lr) * float(np.sqrt(1.0 - BETA2 ** t) / (1.0 - BETA1 ** t))

    def _start_reduce(self, lo, hi):
        """Start the all-reduce of flat-gradient elements [lo, hi) in the wire format `grad_comm` (bf16: the slice is
        cast into the bf16 wire buffer first -- one more pass over it, for half the bytes on NVLink)."""
        if hi <= lo:
            return
        if self.g16 is not None:
            ops.cast_bf16(self.grads[lo:hi], self.g16[lo:hi])
            self.allreduce.start(self.g16[lo:hi])
        else:
            self.allreduce.start(self.grads[lo:hi])

    def _reduce_and_adam(self, lr_t):
        if self.allreduce is not None:
            self._start_reduce(self._reduced_upto, self.n_flat)
            self.allreduce.finish()
            self._reduced_upto = 0
        ops.adam(self.params, self.grads, self.adam_m, self.adam_v, lr_t, BETA1, BETA2, EPS, 1.0 / self.world,
                 w_hi=self.w_hi, w_lo=self.w_lo, lr_ptr=self.step_scalars[0:1],
                 g_bf16=self.g16 if self.allreduce is not None else None)

    @_on_device
    def adam_step(self, lr):
        """TF-form Adam over the flat buffer (one launch), then global_step += 1 (fcn8s_tensorflow.py:256-257)."""
        t = self.global_step + 1
        lr_t = self._lr_t(lr, t)
        ops.set_step_scalars(self.step_scalars, lr_t, 0)
        self._reduce_and_adam(lr_t)
        self.global_step = t
        self._packed_dirty = True

    def _step_body(self, images, labels, keep_prob, l2_rate):
        """Everything one training `sess.run` does on the device, reading lr_t / seed from `step_scalars`: this is the
        unit that is captured into a CUDA graph."""
        self.loss_and_backward(images, labels, keep_prob, l2_rate, 0, _scalars_set=True)
        self._reduce_and_adam(0.0)
        self._packed_dirty = True
        self.repack()     # derived operands of the updated parameters, ready for the next step's forward

    @_on_device
    def train_step(self, images, labels, lr, keep_prob=0.5, l2_rate=0.0, seed=None):
        """One `sess.run([train_op, total_loss, global_step])` (fcn8s_tensorflow.py:565-572).
        Returns the device tensor loss_buf; total_loss = loss_buf[0] / (N*H*W) + loss_buf[1] (see `loss_value`).

        After two eager steps per (shape, keep_prob, l2_rate) the whole step (~250 kernel launches and the gradient
        all-reduce) is captured into a CUDA graph and replayed: the batch is copied into fixed input buffers and the
        two per-step scalars (lr_t, dropout seed) are written to device memory by a 1-thread kernel, so the host's
        cost per step is three launches.  FCN8_GRAPHS=0 (or an installed ops.TIMER) keeps the eager path."""
        if seed is None:
            # data parallel: every rank draws its own dropout masks (seed = f(step, rank)), like independent replicas
            seed = self.global_step * self.world + self.rank
        t = self.global_step + 1
        lr_t = self._lr_t(lr, t)
        key = (tuple(images.shape), float(keep_prob), float(l2_rate))
        graphs = self.use_graphs and ops.TIMER is None
        if graphs:
            A = self._arena(images.shape[0], images.shape[1], images.shape[2])
            xin = self._buf(A, "x_in", images.shape, torch.uint8)
            yin = self._buf(A, "y_in", labels.shape, torch.uint8)
            xin.copy_(images)
            yin.copy_(labels.view(torch.uint8))
            images, labels = xin, yin
        ops.set_step_scalars(self.step_scalars, lr_t, seed)
        entry = self._graphs.get(key) if graphs else None
        if entry is not None:
            if self._packed_dirty or (self.hwio and self._shadow_dirty):
                self.repack()     # weights were replaced (load_weights) since the last step
            entry[0].replay()
            self.graph_launches += entry[1]
        elif graphs and self._warm.get(key, 0) >= 2:
            torch.cuda.synchronize(self.device)
            l0 = self.lib.fcn8_launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._step_body(images, labels, keep_prob, l2_rate)
            n = int(self.lib.fcn8_launch_count() - l0)
            self._graphs[key] = (g, n)
            g.replay()
            self.graph_launches += n
        else:
            self._warm[key] = self._warm.get(key, 0) + 1
            self._step_body(images, labels, keep_prob, l2_rate)
        self.global_step = t
        return self.loss_buf

    def loss_value(self, images_shape):
        N, H, W = images_shape[0], images_shape[1], images_shape[2]
        v = self.loss_buf.tolist()
        return v[0] / float(N * H * W) + v[1]

    # ------------------------------------------------------------------ predictor / evaluation
    @_on_device
    def predict(self, images, argmax=True):
        """fcn8s_tensorflow.py:743-770: argmax int64 [N,H,W] or softmax fp32 [N,H,W,C], keep_prob = 1."""
        N, H, W, _ = images.shape
        self.forward(images, 1.0, 0, train=False)
        zp = self._arena(N, H, W)["logits_p"]
        if argmax:
            out = torch.empty((N, H, W), dtype=torch.int64, device=self.device)
            ops.softmax_xent(zp, argmax=out, pad=4, num_classes=self.C)
        else:
            out = torch.empty((N, H, W, self.C), dtype=torch.float32, device=self.device)
            ops.softmax_xent(zp, softmax=out, pad=4, num_classes=self.C)
        return out

    @_on_device
    def eval_step(self, images, labels, conf, l2_rate=0.0):
        """One metric update (fcn8s_tensorflow.py:685-689): forward at keep_prob 1, total_loss, argmax, confusion
        matrix accumulate (conf: int64 [C,C] device tensor, conf[label, prediction])."""
        N, H, W, _ = images.shape
        self.forward(images, 1.0, 0, train=False)
        zp = self._arena(N, H, W)["logits_p"]
        self.loss_buf.zero_()
        am = torch.empty((N, H, W), dtype=torch.int64, device=self.device)
        lab = labels.view(torch.uint8)
        ops.softmax_xent(zp, lab, self.loss_buf[0:1], argmax=am, pad=4, num_classes=self.C)
        if l2_rate != 0.0:
            for kname in DECODER_KERNELS:
                ops.l2_reg(self.view(kname).reshape(-1), None, self.loss_buf[1:2], l2_rate)
        ops.confusion_matrix(am, lab, conf)
        return self.loss_buf
