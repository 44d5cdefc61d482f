"""FCN-8s forward / backward / Adam engine on libfcn8s_sm100.so.

This is the replacement for what `tf.Session.run` executes in the reference: the external VGG-16 encoder graph
(fcn8s_tensorflow.py:127-152), `_build_decoder` (:154-237), `_build_optimizer` (:239-259) and `_build_predictor`
(:261-271).  torch owns the device memory (flat parameter / gradient / Adam buffers, a per-shape activation arena) and
the stream; every arithmetic operation is a hand-written sm_100a kernel reached through the C ABI
(include/fcn8s_b200.h).  There is no CPU path: constructing an Engine without a CUDA device or without the built
library raises.

Precision modes: see `Engine`.  The decoder (score heads, transposed convolutions) runs on the same tcgen05 GEMM
kernels over bf16 hi / lo planes; the skip adds, the softmax-cross-entropy, its gradient, softmax / argmax and the
confusion matrix live in the epilogues of the transposed-convolution kernels that precede them.
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
# (packed-operand key, TF variable scope, encoder tensor, input channels, skip scale) of the 1x1 score heads (:171-200)
HEADS = [("h3", "pool3_1x1", "pool3", 256, POOL3_SCALE), ("h4", "pool4_1x1", "pool4", 512, POOL4_SCALE),
         ("h7", "fc7_1x1", "fc7", 4096, 1.0)]
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
    """Precision modes (`precision=`), both with fp32 accumulation in TMEM, fp32 master weights / gradients / Adam:

      "fp32"    fp32-equivalent: activations and weights are bf16 hi/lo PAIRS (FCN8_BF16X2: v ~ hi + lo, 16-17
                mantissa bits) and every GEMM forms the error-compensated hi*lo + lo*hi + hi*hi on the bf16 tensor
                cores (3 MMAs per product, the low-order products first).  Meets the 1e-4 logit tolerance.
      "bf16"    activations + tensor-core operands bf16, one MMA per product.
    The GEMMs read a bf16 (hi/lo) shadow of the flat parameter buffer in its TF layout directly (refreshed by the Adam
    kernel), so there is no per-step weight packing of the encoder."""

    MODES = ("bf16", "fp32")

    def __init__(self, num_classes, precision="bf16", device=None, backward_terms=3, grad_comm=None):
        """backward_terms (precision "fp32" only): bf16 products per algorithmic product in the BACKWARD GEMMs (dgrad and
        filter gradients): 3 = like the forward pass (default); 2 / 1 are the measured, non-default "fp32 forward /
        reduced backward" modes (one operand, or both, rounded to bf16).
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
        self.C = num_classes
        self.precision = precision
        self.pair = precision == "fp32"                      # bf16 hi/lo pair activations
        self.nseg = 3 if self.pair else 1                    # MMAs per product, forward
        self.bt = int(backward_terms) if self.pair else 1    # MMAs per product, backward
        self.grad_comm = grad_comm or os.environ.get("FCN8_GRAD_COMM") or ("bf16" if precision == "bf16" else "fp32")
        if self.grad_comm not in ("fp32", "bf16"):
            raise ValueError("grad_comm must be 'fp32' or 'bf16'")
        self.rank = 0
        self.g16 = None          # bf16 wire copy of the flat gradient (allocated by dist.attach when grad_comm == bf16)
        self.dt = ops.BF16
        self.tdt = torch.bfloat16
        self.cm = 2 if self.pair else 1                      # stored channels per logical channel
        self.kp = 64                                         # padded im2col width of conv1_1
        self.conv_algo = int(os.environ.get("FCN8_CONV_ALGO", "0"))   # diagnostics: 1 = per-tap kernels only
        self.fuse_pool = os.environ.get("FCN8_FUSE_POOL", "1") != "0"
        self.conv1_direct = os.environ.get("FCN8_CONV1_IM2COL", "0") == "0"   # conv1_1 from the uint8 image (conv1.cu)
        self.layers = encoder_layers()
        self.layout, self.n_flat = flat_layout(num_classes)
        self.bias_block = self.layout["fc7/biases"][0]       # encoder biases: [bias_block, n_flat)
        z = dict(dtype=torch.float32, device=self.device)
        self.params = torch.zeros(self.n_flat, **z)
        self.grads = torch.zeros(self.n_flat, **z)
        self.adam_m = torch.zeros(self.n_flat, **z)
        self.adam_v = torch.zeros(self.n_flat, **z)
        self.w_hi = torch.zeros(self.n_flat, dtype=torch.bfloat16, device=self.device)
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
        self.mid_elems = self.layout["conv3_3/filter"][0]    # ... | conv5_x W | conv4_x W: final after conv4_1's wgrad
        self._wire_stream = None   # side stream of the bf16 wire cast (created on first use)
        self.sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.dp_reserve_sms = int(os.environ.get("FCN8_DP_RESERVE_SMS", "0"))
        # ... for this many encoder layers of the backward pass after the collective was started (0 = until the end)
        self.dp_reserve_layers = int(os.environ.get("FCN8_DP_RESERVE_LAYERS", "0"))

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
        """Derived tensor-core operands of the parameters: the encoder reads the bf16 shadow written by Adam; only conv1_1
        (27 -> 64 im2col columns), the three score heads (classes padded to 64 columns) and the phase-GEMM operands
        of the three transposed convolutions -- 0.2 % of the parameters -- are packed."""
        if self._shadow_dirty:
            ops.shadow_weights(self.params, self.w_hi, self.w_lo)
            self._shadow_dirty = False
        old = self.packed.get("conv1_1")   # refilled in place: a captured CUDA graph keeps reading the same buffers
        # runs as a 1x1 conv over the 27-column im2col (padded to kp) built by the feed kernel
        self.packed["conv1_1"] = ops.pack_weights(self.view("conv1_1/filter"), 1, 27, 64, 0, self.dt, cin_pad=self.kp,
                                                  split=self.pair, out=old)
        # conv1_2 forward on CTA pairs (conv_halo_pair_kernel) needs K-major weights: each CTA of a pair holds 32 of the
        # 64 output channels, too narrow for the MN-major operand read from the TF layout (36 864 elements)
        self.packed["conv1_2"] = ops.pack_weights(self.view("conv1_2/filter"), 3, 64, 64, 0, self.dt, split=self.pair,
                                                  out=self.packed.get("conv1_2"))
        for key, base, _, cin, _ in HEADS:
            self.packed[key] = ops.head_pack(self.view(base + "/kernel").view(cin, self.C), self.view(base + "/bias"),
                                             split=self.pair, out=self.packed.get(key))
        for key, base, stride in UPSCORE_STAGES:
            self.packed[key] = ops.deconv_pack(self.view(base + "/kernel"), self.view(base + "/bias"), stride,
                                               split=self.pair, out=self.packed.get(key))
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

    def _buf(self, arena, name, shape, dtype, zero=False):
        t = arena.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            arena[name] = t
        return t

    def _act(self, arena, name, N, h, w, c):
        """Encoder activation buffer [N,h,w,c] in the mode's storage format ([N,h,w,2c] bf16 for hi/lo pairs)."""
        return self._buf(arena, name, (N, h, w, c * self.cm), self.tdt)

    def _planes(self, arena, name, N, h, w):
        """Decoder activation: a bf16 hi/lo pair tensor [N,h,w,128] (64 hi | 64 lo channels, classes zero-padded);
        returns its (hi, lo) plane views."""
        return ops.halves(self._buf(arena, name, (N, h, w, 128), torch.bfloat16), 64)

    def _padded(self, arena, name, N, h, w, stride):
        """Zero-bordered padded blocked gradient planes of a stride-s stage (allocated and zeroed once: the kernels
        only ever write the interior and write zeros beyond the class count)."""
        p = arena.get(name)
        if p is None:
            p = arena[name] = ops.padded_alloc(N, h, w, stride, self.device)
        return p

    def _wview(self, name):
        """(hi, lo) bf16 shadow views of an encoder weight tensor in TF layout."""
        return self.view(_wname(name), self.w_hi), (self.view(_wname(name), self.w_lo) if self.pair else None)

    def _conv(self, name, x, k, cout, out, bias, flags, **kw):
        """Encoder convolution of layer `name` (forward direction) in the mode's operand format."""
        if name == "conv1_1":
            wp, wlo = self.packed[name]
            return ops.conv_gemm(x, wp, cout, 1, bias=bias, flags=flags, out=out, wp_lo=wlo, pair=self.pair, **kw)
        if name in self.packed:     # conv1_2 in the fp32-equivalent mode: packed K-major forward operand
            wp, wlo = self.packed[name]
            return ops.conv_gemm(x, wp, cout, k, bias=bias, flags=flags, out=out, wp_lo=wlo, pair=self.pair, w_mode=0,
                                 algo=self.conv_algo, **kw)
        wh, wl = self._wview(name)
        return ops.conv_gemm(x, wh, cout, k, bias=bias, flags=flags, out=out, wp_lo=wl, pair=self.pair, w_mode=1,
                             algo=self.conv_algo, **kw)

    # ------------------------------------------------------------------ forward
    def _features(self, images, keep_prob, seed, train, _scalars_set=False):
        """images -> f3 planes (the input of upscore8).  Everything below the loss / predictor: feed pre-processing,
        VGG-16 encoder (the 2x2 max-pools in the epilogue of the convolution that precedes them), score heads,
        upscore2 (+ pool4 skip) and upscore_pool4 (+ pool3 skip)."""
        if self._packed_dirty or self._shadow_dirty:
            self.repack()
        if train and keep_prob < 1.0 and not _scalars_set:
            ops.set_step_scalars(self.step_scalars, 0.0, seed)
        N, H, W, _ = images.shape
        A = self._arena(N, H, W)
        A["_shape"] = (N, H, W)
        pair = self.pair
        A["images"] = images
        x = None
        if not self.conv1_direct:   # the global-im2col variant of conv1_1 (kept for A/B measurements)
            x = self._act(A, "im2col", N, H, W, self.kp)
            p = capi.PreprocessParams(capi.ptr(images), capi.ptr(x), N, H, W, ops.BF16X2 if pair else self.dt)
            capi.check(self.lib.fcn8_preprocess_im2col(ops.C.byref(p), ops._stream()))
        h, w = H, W
        li = 0
        for b, cout, n in VGG_BLOCKS:
            for i in range(1, n + 1):
                name, k, cin, _ = self.layers[li]
                li += 1
                kw = {}
                pooled = None
                if i == n and self.fuse_pool:
                    # last conv of the block: its epilogue also emits the block's max-pool; the full-resolution tensor
                    # is only stored when the backward pass will need it
                    pooled = self._act(A, "pool%d" % b, N, h // 2, w // 2, cout)
                    kw = dict(pool_out=pooled, store_out=train)
                out = self._act(A, name, N, h, w, cout) if (train or pooled is None) else None
                if name == "conv1_1" and self.conv1_direct:
                    ops.conv1_fwd(images, self.packed[name], self.view(name + "/biases"), out, pair=pair)
                else:
                    self._conv(name, x, k, cout, out, self.view(name + "/biases"), ops.EPI_BIAS | ops.EPI_RELU, **kw)
                x = out
            h, w = h // 2, w // 2
            if not self.fuse_pool:
                pooled = self._act(A, "pool%d" % b, N, h, w, cout)
                ops.maxpool_fwd(x, out=pooled, pair=pair)
            x = pooled
        drop = train and keep_prob < 1.0
        for name, k, cin, cout in self.layers[-2:]:
            out = self._act(A, name, N, h, w, cout)
            flags = ops.EPI_BIAS | ops.EPI_RELU | (ops.EPI_DROPOUT if drop else 0)
            self._conv(name, x, k, cout, out, self.view(name + "/biases"), flags,
                       keep_prob=keep_prob if drop else 1.0, seed=1 if name == "fc7" else 0,
                       seed_ptr=self.step_scalars[1:2] if drop else None)
            x = out
        # score heads (fcn8s_tensorflow.py:171-200): 1x1 convolutions with the class dimension padded to 64 columns, the
        # 1e-4 / 1e-2 skip scales folded into the accumulator scale
        S = {}
        for key, base, src, cin, scale in HEADS:
            t = A[src]
            n_, hh, ww, _ = t.shape
            S[key] = self._buf(A, "s_" + key, (n_, hh, ww, 128), torch.bfloat16)
            pk = self.packed[key]
            ops.conv_gemm(t, pk["w"], 64, 1, bias=pk["bias64"], flags=ops.EPI_BIAS, out=S[key], wp_lo=pk["w_lo"],
                          pair=pair, out_pair=True, w_mode=1, out_scale=scale, nseg=self.nseg, tag="head_gemm")
        C = self.C
        # upscore2 + pool4 skip, upscore_pool4 + pool3 skip (:204-224): the tf.add is the residual operand of the epilogue
        f4 = self._planes(A, "f4", N, H // 16, W // 16)
        ops.deconv_fwd(ops.halves(S["h7"], 64), self.packed["up2"], C, 2, f4, skip=ops.halves(S["h4"], 64),
                       nseg=self.nseg)
        f3 = self._planes(A, "f3", N, H // 8, W // 8)
        ops.deconv_fwd(f4, self.packed["up4"], C, 2, f3, skip=ops.halves(S["h3"], 64), nseg=self.nseg)
        return A, f3

    @_on_device
    def forward(self, images, keep_prob=1.0, seed=0, train=False, _scalars_set=False):
        """images: uint8 CUDA tensor [N,H,W,3] (RGB). Returns fp32 logits [N,H,W,C] (an arena buffer).
        The dropout seed is read by the kernels from `step_scalars` (set here unless the caller already did)."""
        N, H, W, _ = images.shape
        A, f3 = self._features(images, keep_prob, seed, train, _scalars_set)
        logits = self._buf(A, "logits", (N, H, W, self.C), torch.float32)
        ops.deconv_loss(f3, self.packed["up8"], self.C, nseg=self.nseg, logits=logits)
        return logits

    @staticmethod
    def dropout_seed(seed, layer):
        """Seed of the counter-based dropout mask of `layer` for step seed `seed` (host replica: rng.py); the kernels
        form the same value from the device scalar: *seed_ptr * 2 + (layer == fc7)."""
        return (int(seed) * 2 + (1 if layer == "fc7" else 0)) & 0xFFFFFFFF

    # ------------------------------------------------------------------ loss + backward
    @_on_device
    def loss_and_backward(self, images, labels, keep_prob=1.0, l2_rate=0.0, seed=0, _scalars_set=False,
                          store_logits=True):
        """Forward + backward of optimizer/total_loss (fcn8s_tensorflow.py:250-257) into self.grads.
        labels: uint8/bool CUDA tensor [N,H,W,C] one-hot. Returns the device scalar pair loss_buf (CE sum, L2).
        store_logits=False (the training step): the logits never leave the upscore8 kernel's epilogue."""
        N, H, W, _ = images.shape
        C = self.C
        pair = self.pair
        A, f3 = self._features(images, keep_prob, seed, True, _scalars_set)
        G = self.grads
        bt = self.bt
        self.loss_buf.zero_()
        G[self.bias_block:].zero_()   # encoder bias gradients are accumulated by fused epilogues (atomics)
        gv = lambda n: self.view(n, G)   # noqa: E731
        dc8, dc4, dc2, db7 = (gv("fc7_pool4_pool3_conv2d_trans/bias"), gv("fc7_pool4_conv2d_trans/bias"),
                              gv("fc7_conv2d_trans/bias"), gv("fc7_1x1/bias"))
        for t in (dc8, dc4, dc2, db7):
            t.zero_()
        # upscore8 + softmax-CE + its gradient in one kernel (:226-235, :253): dz goes straight into the padded blocked
        # planes the gradient GEMMs read; dc8 = class sums of dz
        dz8 = self._padded(A, "dz8", N, H // 8, W // 8, 8)
        logits = self._buf(A, "logits", (N, H, W, C), torch.float32) if store_logits else None
        ops.deconv_loss(f3, self.packed["up8"], C, nseg=self.nseg, labels=labels.view(torch.uint8),
                        loss_sum=self.loss_buf[0:1], dz_out=dz8 if pair else (dz8[0], None), dbias=dc8,
                        grad_scale=1.0 / (N * H * W), logits=logits)
        # decoder backward: every stage's input gradient lands in the interior of the next
        # stage's padded planes (the skip adds fan the gradient out unchanged: the same planes feed the score heads)
        # and its column sums are the bias gradients of the two layers whose outputs were added there
        dz4 = self._padded(A, "dz4", N, H // 16, W // 16, 2)      # gradient of f3 = output gradient of upscore_pool4
        dz2 = self._padded(A, "dz2", N, H // 32, W // 32, 2)      # gradient of f4 = output gradient of upscore2
        df3, df4 = ops.padded_interior(dz4, 2), ops.padded_interior(dz2, 2)
        ds7 = self._planes(A, "ds7", N, H // 32, W // 32)
        S = {k: ops.halves(A["s_" + k], 64) for k in ("h3", "h4", "h7")}
        ops.deconv_dw(f3, dz8, C, 8, gv("fc7_pool4_pool3_conv2d_trans/kernel"), nseg=bt)
        ops.deconv_dx(dz8, self.packed["up8"], C, 8, df3, nseg=bt, colsum=dc4)
        ops.deconv_dw(self._planes(A, "f4", N, H // 16, W // 16), dz4, C, 2, gv("fc7_pool4_conv2d_trans/kernel"), nseg=bt)
        ops.deconv_dx(dz4, self.packed["up4"], C, 2, df4, nseg=bt, colsum=dc2)
        ops.deconv_dw(S["h7"], dz2, C, 2, gv("fc7_conv2d_trans/kernel"), nseg=bt)
        ops.deconv_dx(dz2, self.packed["up2"], C, 2, ds7, nseg=bt, colsum=db7)
        gv("pool3_1x1/bias").copy_(dc4)     # f3 = upscore_pool4(f4) + c4 + s3: both biases see the same gradient
        gv("pool4_1x1/bias").copy_(dc2)
        inv_keep = 1.0 / keep_prob if keep_prob < 1.0 else 1.0
        # score heads: dK = scale * x^T ds, dx = scale * ds K^T (for fc7 with its ReLU / dropout mask)
        dhead = {}
        for key, base, src, cin, scale in HEADS:
            ds = {"h3": df3, "h4": df4, "h7": ds7}[key]
            x = A[src]
            pk = self.packed[key]
            ops.wgrad_gemm(x, ds[0], 1, gv(base + "/kernel").view(cin, C), dy_lo=ds[1], pair=pair, dy_pair=False,
                           nseg=bt, out_cols=C, out_scale=scale)
            dx = self._buf(A, "d_%s_head" % src, x.shape, self.tdt)
            mk = dict(flags=ops.EPI_MASK, mask_src=x, mask_scale=inv_keep) if key == "h7" else {}
            ops.conv_gemm(ds[0], pk["w"], cin, 1, x_lo=ds[1] if bt == 3 else None, wp_lo=pk["w_lo"], w_mode=2, out=dx,
                          out_pair=pair, out_scale=scale, nseg=bt, cin=64, tag="head_gemm", **mk)
            dhead[src] = dx
        dy = dhead["fc7"]
        if l2_rate != 0.0:
            for kname in DECODER_KERNELS:
                ops.l2_reg(self.view(kname).reshape(-1), self.view(kname, G).reshape(-1), self.loss_buf[1:2], l2_rate)
        # encoder backward
        reserve_left = None
        for li in range(len(self.layers) - 1, -1, -1):
            name, k, cin, cout = self.layers[li]
            x_in = self._layer_input(A, li)
            gw = self.view(_wname(name), G)
            if name == "fc7":   # every other bias gradient is fused into the kernel that produces that layer's dY
                ops.bias_grad(dy, self.view(name + "/biases", G), pair=pair)
            if name == "conv1_1":
                if self.conv1_direct:
                    ops.conv1_wgrad(A["images"], dy, gw.view(27, cout), pair=pair)
                else:
                    ops.wgrad_gemm(x_in, dy, 1, gw.view(27, cout), rows_valid=27, pair=pair, nseg=bt)
                break
            ops.wgrad_gemm(x_in, dy, k, gw.view(k * k * cin, cout), pair=pair, nseg=bt)
            if name == "fc6" and self.allreduce is not None and getattr(self.allreduce, "overlap", False):
                # decoder, fc7 and fc6 gradients (89 % of the buffer) are final: reduce them under the conv backward
                self._start_reduce(0, self.head_elems)
                self._reduced_upto = self.head_elems
                if self.dp_reserve_sms > 0:   # leave SMs to the collective's CTAs while it runs under the backward
                    self.lib.fcn8_set_sm_limit(self.sm_count - self.dp_reserve_sms)
                    reserve_left = self.dp_reserve_layers if self.dp_reserve_layers > 0 else 1 << 30
            elif self.dp_reserve_sms > 0 and reserve_left is not None:
                reserve_left -= 1
                if reserve_left <= 0:
                    self.lib.fcn8_set_sm_limit(0)
                    reserve_left = None
            if name == "conv4_1" and self.allreduce is not None and getattr(self.allreduce, "overlap", False) \
                    and self._reduced_upto == self.head_elems:
                # conv5_x / conv4_x filter gradients (88 % of what is left) are final: second chunk, so that only the
                # conv3..conv1 filters and the bias block (1.7 M elements) remain for the exposed tail before Adam
                self._start_reduce(self.head_elems, self.mid_elems)
                self._reduced_upto = self.mid_elems
            dx = self._buf(A, "dx_" + name, x_in.shape, self.tdt)
            prev_name = self.layers[li - 1][0]
            prev_db = self.view(prev_name + "/biases", G)
            wh, wl = self._wview(name)
            wkw = dict(wp_lo=wl, pair=pair, w_mode=2, nseg=bt, algo=self.conv_algo)
            if self._input_is_pool(li):
                # x_in is a pool output: no ReLU mask here (the pool backward applies it); add the score-head
                # gradient at pool3 / pool4 (AddN of the two consumers)
                pool_idx = self._pool_index(li)
                res = dhead["pool3"] if pool_idx == 3 else (dhead["pool4"] if pool_idx == 4 else None)
                ops.conv_gemm(dy, wh, cin, k, flags=(ops.EPI_RESIDUAL if res is not None else 0), residual=res, out=dx,
                              **wkw)
                src = A[prev_name]  # pre-pool activation (post-ReLU)
                dpre = self._buf(A, "dpre_" + prev_name, src.shape, self.tdt)
                ops.maxpool_bwd(src, dx, out=dpre, pair=pair, db=prev_db)
                dy = dpre
            else:
                # ReLU (and for fc6 -> dropout) backward of the producer fused as an epilogue mask on its output
                scale = inv_keep if prev_name == "fc6" else 1.0
                ops.conv_gemm(dy, wh, cin, k, flags=ops.EPI_MASK, mask_src=x_in, mask_scale=scale, out=dx,
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
            return A.get("im2col")
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
            # The cast is a 1.4 GB pass for the head chunk: it runs on a side stream, next to the (tensor-bound) conv
            # backward kernels that follow on the compute stream, and the collective is ordered after it -- the
            # compute stream only meets both again in allreduce.finish().
            cur = torch.cuda.current_stream(self.device)
            if self._wire_stream is None:
                self._wire_stream = torch.cuda.Stream(device=self.device)
            self._wire_stream.wait_stream(cur)
            with torch.cuda.stream(self._wire_stream):
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
        self.loss_and_backward(images, labels, keep_prob, l2_rate, 0, _scalars_set=True, store_logits=False)
        self._reduce_and_adam(0.0)
        self._packed_dirty = True
        self.repack()     # derived operands of the updated parameters, ready for the next step's forward

    @_on_device
    def train_step(self, images, labels, lr, keep_prob=0.5, l2_rate=0.0, seed=None):
        """One `sess.run([train_op, total_loss, global_step])` (fcn8s_tensorflow.py:565-572).
        Returns the device tensor loss_buf; total_loss = loss_buf[0] / (N*H*W) + loss_buf[1] (see `loss_value`).

        After two eager steps per (shape, keep_prob, l2_rate) the whole step (~110 kernel launches and the gradient
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
            if self._packed_dirty or self._shadow_dirty:
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
    def predict(self, images, argmax=True, compact=False):
        """fcn8s_tensorflow.py:743-770: argmax int64 [N,H,W] or softmax fp32 [N,H,W,C], keep_prob = 1 -- computed in
        the epilogue of the upscore8 kernel; the logits are never written.  compact=True: the class map as uint8
        [N,H,W] (one byte per pixel for the trip to the host; the caller widens it to tf.argmax's int64)."""
        N, H, W, _ = images.shape
        A, f3 = self._features(images, 1.0, 0, False)
        if argmax and compact:
            out = torch.empty((N, H, W), dtype=torch.uint8, device=self.device)
            ops.deconv_loss(f3, self.packed["up8"], self.C, nseg=self.nseg, argmax_u8=out)
        elif argmax:
            out = torch.empty((N, H, W), dtype=torch.int64, device=self.device)
            ops.deconv_loss(f3, self.packed["up8"], self.C, nseg=self.nseg, argmax=out)
        else:
            out = torch.empty((N, H, W, self.C), dtype=torch.float32, device=self.device)
            ops.deconv_loss(f3, self.packed["up8"], self.C, nseg=self.nseg, softmax=out)
        return out

    @_on_device
    def eval_step(self, images, labels, conf, l2_rate=0.0):
        """One metric update (fcn8s_tensorflow.py:685-689): forward at keep_prob 1, total_loss, argmax and the
        confusion-matrix accumulate (conf: int64 [C,C] device tensor, conf[label, prediction]) in the epilogue of the
        upscore8 kernel."""
        A, f3 = self._features(images, 1.0, 0, False)
        self.loss_buf.zero_()
        ops.deconv_loss(f3, self.packed["up8"], self.C, nseg=self.nseg, labels=labels.view(torch.uint8),
                        loss_sum=self.loss_buf[0:1], conf=conf)
        if l2_rate != 0.0:
            for kname in DECODER_KERNELS:
                ops.l2_reg(self.view(kname).reshape(-1), None, self.loss_buf[1:2], l2_rate)
        return self.loss_buf
