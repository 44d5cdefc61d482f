"""Host-side replica of the counter-based dropout RNG used by the CUDA epilogues (csrc/conv_gemm.cuh: mix32 /
dropout_keep). The reference's dropout uses TensorFlow's unseeded RNG (keep_prob fed at fcn8s_tensorflow.py:561), so
bit-parity of masks with TF is impossible by construction; parity tests instead regenerate THIS mask on the host and
inject it into the oracle."""
import numpy as np
import torch


def mix32(seed, idx):
    """idx: uint64 numpy array of element indices -> uint32 hash (same arithmetic as the device function)."""
    m = np.uint64(0xFFFFFFFF)
    idx = idx.astype(np.uint64)
    seed = np.uint64(seed & 0xFFFFFFFF)
    x = (idx & m) ^ (((idx >> np.uint64(32)) * np.uint64(0x9E3779B9)) & m) ^ seed
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & m
    x ^= x >> np.uint64(16)
    x = (x + ((seed * np.uint64(0x9E3779B9)) & m)) & m
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x2C1B3C6D)) & m
    x ^= x >> np.uint64(12)
    return x.astype(np.uint32)


def keep_threshold(keep_prob):
    return int(float(np.float32(keep_prob)) * 16777216.0) & 0xFFFFFFFF


def dropout_keep_mask(seed, n, keep_prob):
    """Boolean torch tensor [n]: True where the element at flat index i is kept."""
    h = mix32(seed, np.arange(n, dtype=np.uint64))
    return torch.from_numpy((h >> np.uint32(8)) < np.uint32(keep_threshold(keep_prob)))
