"""Per-kernel GPU parity through the C ABI against fp64 torch references of the same op (scripts/bringup.py holds the
cases so the same code serves bring-up diagnostics): tcgen05 implicit-GEMM conv fprop / dgrad / wgrad for every tile
width, split-K, 1x1 / 3x3 / 7x7 geometries, ragged tiles, bf16 / tf32 / 3xTF32; epilogues with explicit masks; the
HBM-bound kernels (feed, pool, bias-grad, Adam, L2); decoder, loss, predictor and confusion-matrix kernels.
Tolerances are written next to each case in scripts/bringup.py (bf16 1e-2, tf32 2e-3, 3xTF32 2e-6, fp32 CUDA-core
kernels 1e-5, integer / max-pool results exact)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("bringup", os.path.join(ROOT, "scripts", "bringup.py"))
bringup = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(bringup)


@pytest.mark.parametrize("name", list(bringup.CASES))
def test_kernel_case(cuda_device, name):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    assert bringup.CASES[name](), name
    torch.cuda.synchronize()


def test_confusion_matrix_matches_reference_c_loop(cuda_device):
    """Device confusion-matrix kernel vs the reference's own native loop
    (cityscapesscripts/evaluation/addToConfusionMatrix_impl.c:3-16) compiled into oracle/_ref by oracle/Makefile."""
    import ctypes
    so = os.path.join(ROOT, "oracle", "_ref", "libaddtoconfusionmatrix.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    from fcn8s_tensorflow_b200 import ops
    ref = ctypes.CDLL(so)
    rng = np.random.default_rng(0)
    C, Hh, Ww = 20, 97, 131
    gt = rng.integers(0, C, (Hh, Ww)).astype(np.uint8)
    pr = rng.integers(0, C, (Hh, Ww)).astype(np.uint8)
    cm_ref = np.zeros((C, C), np.uint64)
    ref.addToConfusionMatrix(pr.ctypes.data_as(ctypes.c_void_p), gt.ctypes.data_as(ctypes.c_void_p),
                             ctypes.c_uint(Ww), ctypes.c_uint(Hh), cm_ref.ctypes.data_as(ctypes.c_void_p),
                             ctypes.c_uint(C))
    onehot = torch.from_numpy(np.eye(C, dtype=np.uint8)[gt]).to(cuda_device)
    pred = torch.from_numpy(pr.astype(np.int64)).to(cuda_device)
    conf = torch.zeros((C, C), dtype=torch.int64, device=cuda_device)
    ops.confusion_matrix(pred, onehot, conf)
    assert np.array_equal(conf.cpu().numpy().astype(np.uint64), cm_ref)
