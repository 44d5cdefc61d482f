"""TensorBoard variable summaries of the reference (`_build_summary_ops`, fcn8s_tensorflow.py:324-369, with
helpers/tf_variable_summaries.py:3-20): per variable the scalars mean / stddev / max / min and a histogram, for the six
decoder layers and fc7, fc6, conv4_3, conv3_3 (kernel and bias each), written next to `total_loss` and `learning_rate`
every `summaries_frequency` steps.

Off the hot path, but fc6 alone is 103 M values: everything is reduced where the parameters live (torch ops on the flat
parameter buffer's views) and only the 4 scalars and the bucket counts cross to the host.  The histogram uses
TensorFlow's default bucket limits (tensorflow/core/lib/histogram/histogram.cc: +-1e-12 * 1.1^k up to 1e20, a zero
bucket, DBL_MAX) and its `upper_bound` bucket rule, so TensorBoard shows the same distribution the reference's
`tf.summary.histogram` would."""
import numpy as np
import torch

# (TF variable name, summary scope) in the reference's order, fcn8s_tensorflow.py:331-350
SUMMARY_VARIABLES = [
    ("pool3_1x1/kernel", "pool3_1x1/kernel"), ("pool3_1x1/bias", "pool3_1x1/bias"),
    ("pool4_1x1/kernel", "pool4_1x1/kernel"), ("pool4_1x1/bias", "pool4_1x1/bias"),
    ("fc7_1x1/kernel", "fc7_1x1/kernel"), ("fc7_1x1/bias", "fc7_1x1/bias"),
    ("fc7_conv2d_trans/kernel", "fc7_conv2d_trans/kernel"), ("fc7_conv2d_trans/bias", "fc7_conv2d_trans/bias"),
    ("fc7_pool4_conv2d_trans/kernel", "fc7_pool4_conv2d_trans/kernel"),
    ("fc7_pool4_conv2d_trans/bias", "fc7_pool4_conv2d_trans/bias"),
    ("fc7_pool4_pool3_conv2d_trans/kernel", "fc7_pool4_pool3_conv2d_trans/kernel"),
    ("fc7_pool4_pool3_conv2d_trans/bias", "fc7_pool4_pool3_conv2d_trans/bias"),
    ("fc7/weights", "fc7/kernel"), ("fc7/biases", "fc7/bias"),
    ("fc6/weights", "fc6/kernel"), ("fc6/biases", "fc6/bias"),
    ("conv4_3/filter", "conv4_3/kernel"), ("conv4_3/biases", "conv4_3/bias"),
    ("conv3_3/filter", "conv3_3/kernel"), ("conv3_3/biases", "conv3_3/bias"),
]

_LIMITS = None


def tf_bucket_limits():
    """TensorFlow's default histogram bucket limits (1549 ascending doubles)."""
    global _LIMITS
    if _LIMITS is None:
        pos = []
        v = 1e-12
        while v < 1e20:
            pos.append(v)
            v *= 1.1
        _LIMITS = np.array([-x for x in reversed(pos)] + [0.0] + pos + [np.finfo(np.float64).max], np.float64)
    return _LIMITS


def variable_stats(t):
    """(mean, stddev, max, min) as in add_variable_summaries: population standard deviation around the mean."""
    x = t.reshape(-1).float()
    mean = x.double().mean()
    std = (x.double() - mean).square().mean().sqrt()
    return torch.stack([mean, std, x.max().double(), x.min().double()])


def variable_histogram(t):
    """TensorFlow-bucketed histogram of a tensor, computed on the tensor's device.
    Returns dict(min, max, num, sum, sum_squares, bucket_limits, bucket_counts) with numpy / python values; the
    buckets run from the first to the last non-empty one."""
    x = t.reshape(-1).double()
    limits = torch.from_numpy(tf_bucket_limits()).to(x.device)
    # bucket b holds limits[b-1] <= v < limits[b]  (std::upper_bound in Histogram::Add)
    idx = torch.bucketize(x, limits, right=True).clamp_(max=limits.numel() - 1)
    counts = torch.bincount(idx, minlength=limits.numel())
    head = torch.stack([x.min(), x.max(), x.sum(), x.square().sum()]).cpu().numpy()
    counts = counts.cpu().numpy()
    nz = np.nonzero(counts)[0]
    lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 1)
    return {"min": float(head[0]), "max": float(head[1]), "num": int(x.numel()), "sum": float(head[2]),
            "sum_squares": float(head[3]), "bucket_limits": tf_bucket_limits()[lo:hi].tolist(),
            "bucket_counts": counts[lo:hi].astype(np.float64).tolist()}


def write_variable_summaries(writer, tensors, step):
    """`tensors`: TF variable name -> tensor (any device).  Tags follow the reference's name scopes:
    `<scope>/mean`, `<scope>/stddev`, `<scope>/max`, `<scope>/min`, `<scope>/histogram`."""
    present = [(n, s) for n, s in SUMMARY_VARIABLES if n in tensors]
    if not present:
        return
    stats = torch.stack([variable_stats(tensors[n]) for n, _ in present]).cpu().numpy()   # one device -> host copy
    for (name, scope), row in zip(present, stats):
        for tag, v in zip(("mean", "stddev", "max", "min"), row):
            writer.add_scalar("%s/%s" % (scope, tag), float(v), step)
        h = variable_histogram(tensors[name])
        writer.add_histogram_raw("%s/histogram" % scope, min=h["min"], max=h["max"], num=h["num"], sum=h["sum"],
                                 sum_squares=h["sum_squares"], bucket_limits=h["bucket_limits"],
                                 bucket_counts=h["bucket_counts"], global_step=step)
