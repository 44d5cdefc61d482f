"""First-principles numpy restatement of the TensorFlow-1.x op DEFINITIONS the reference graph relies on -- TEST
INFRASTRUCTURE ONLY (see oracle/fcn8s_oracle.py header; parity unpinned, TF 1.x cannot run here).

TensorFlow (README.md:30, floor 1.0 at fcn8s_tensorflow.py:37, notebook ran 1.3.0) is the un-vendored third-party
module that holds the arithmetic.  These are its published op semantics written as explicit loops, independent of
torch, so that tests/test_oracle.py can check the torch-based oracle (padding conventions, kernel layouts, transposed
convolution as the input-gradient of a strided SAME convolution, Adam with epsilon outside the square root) on tiny
shapes.  Deliberately slow and simple.
"""
import numpy as np


def same_padding(in_size, k, stride):
    """TF 'SAME': out = ceil(in/stride); total pad = max((out-1)*stride + k - in, 0); extra pad goes after."""
    out = -(-in_size // stride)
    total = max((out - 1) * stride + k - in_size, 0)
    return out, total // 2, total - total // 2


def conv2d_same(x, w, stride=1):
    """tf.nn.conv2d(x NHWC, w HWIO, strides=stride, padding='SAME') (cross-correlation)."""
    n, h, wd, cin = x.shape
    kh, kw, _, cout = w.shape
    oh, pt, _ = same_padding(h, kh, stride)
    ow, pl, _ = same_padding(wd, kw, stride)
    y = np.zeros((n, oh, ow, cout), np.float64)
    for i in range(oh):
        for j in range(ow):
            for a in range(kh):
                for b in range(kw):
                    yy, xx = i * stride + a - pt, j * stride + b - pl
                    if 0 <= yy < h and 0 <= xx < wd:
                        y[:, i, j, :] += x[:, yy, xx, :].astype(np.float64) @ w[a, b].astype(np.float64)
    return y


def conv2d_transpose_same(x, t, stride):
    """tf.layers.conv2d_transpose(kernel [kh,kw,Cout,Cin], strides=stride, padding='same') as used at
    fcn8s_tensorflow.py:204-233: by definition the gradient w.r.t. the INPUT of conv2d_same(., filter=t viewed as
    HWIO with I=Cout_T and O=Cin_T, stride), for an input of spatial size stride*h x stride*w."""
    n, h, wd, cin = x.shape
    kh, kw, cout, cin2 = t.shape
    assert cin == cin2
    H, W = h * stride, wd * stride
    oh, pt, _ = same_padding(H, kh, stride)
    ow, pl, _ = same_padding(W, kw, stride)
    assert (oh, ow) == (h, wd)
    y = np.zeros((n, H, W, cout), np.float64)
    # forward conv: z[i,j,ci] = sum_{a,b,co} u[i*s+a-pt, j*s+b-pl, co] * t[a,b,co,ci];  dL/du is the transpose map.
    for i in range(h):
        for j in range(wd):
            for a in range(kh):
                for b in range(kw):
                    yy, xx = i * stride + a - pt, j * stride + b - pl
                    if 0 <= yy < H and 0 <= xx < W:
                        y[:, yy, xx, :] += x[:, i, j, :].astype(np.float64) @ t[a, b].astype(np.float64).T
    return y


def max_pool_2x2_same(x):
    """tf.nn.max_pool(ksize 2, strides 2, 'SAME'): padded positions never win."""
    n, h, wd, c = x.shape
    oh, pt, _ = same_padding(h, 2, 2)
    ow, pl, _ = same_padding(wd, 2, 2)
    y = np.full((n, oh, ow, c), -np.inf, np.float64)
    for i in range(oh):
        for j in range(ow):
            for a in range(2):
                for b in range(2):
                    yy, xx = i * 2 + a - pt, j * 2 + b - pl
                    if 0 <= yy < h and 0 <= xx < wd:
                        y[:, i, j, :] = np.maximum(y[:, i, j, :], x[:, yy, xx, :])
    return y


def softmax_cross_entropy_with_logits(labels, logits):
    """tf.nn.softmax_cross_entropy_with_logits: -sum_c labels_c * log_softmax(logits)_c over the last axis."""
    z = logits.astype(np.float64)
    z = z - z.max(-1, keepdims=True)
    logsm = z - np.log(np.exp(z).sum(-1, keepdims=True))
    return -(labels.astype(np.float64) * logsm).sum(-1)


def adam_apply(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer update for step t (1-based): epsilon is added to sqrt(v), not to the bias-corrected
    sqrt(v_hat)."""
    lr_t = lr * np.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    p = p - lr_t * m / (np.sqrt(v) + eps)
    return p, m, v


def mean_iou(cm):
    """tf.metrics.mean_iou from a confusion matrix cm[label, prediction]."""
    cm = cm.astype(np.float64)
    rows, cols, diag = cm.sum(1), cm.sum(0), np.diag(cm)
    denom = rows + cols - diag
    nvalid = (denom != 0).sum()
    iou = diag / np.where(denom > 0, denom, 1.0)
    return float(iou.sum() / nvalid) if nvalid > 0 else 0.0
