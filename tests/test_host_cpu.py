"""CPU tier for the host side: C-ABI library loads and exports every declared symbol (no compute calls), flat-buffer
layout, dropout RNG replica, loud failure without a GPU, data-parallel host logic over gloo (world size 2)."""
import contextlib
import ctypes
import io
import os
import sys
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from fcn8s_tensorflow_b200 import build
    return build.build()


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "fcn8s_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(fcn8_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(lib_path):
    names = declared_symbols()
    assert len(names) >= 25
    lib = ctypes.CDLL(lib_path)
    for n in names:
        assert hasattr(lib, n), n
    from fcn8s_tensorflow_b200 import _capi
    assert sorted(_capi.EXPORTS) == names
    lib.fcn8_version.restype = ctypes.c_int32
    assert lib.fcn8_version() == 100


def test_header_is_valid_c_and_a_c_client_links(lib_path, tmp_path):
    """include/fcn8s_b200.h is the drop-in boundary: it must compile as plain C (no torch / C++ types in the
    signatures) and a C client must link against the in-tree library and get the documented status codes."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(lib_path)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L", libdir,
                    "-l:" + os.path.basename(lib_path), "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "abi ok" in r.stdout, (r.returncode, r.stdout)


def test_pack_labels_host_code_is_exact_and_rejects_anything_but_one_hot(lib_path):
    """fcn8_pack_labels is HOST code (no CUDA call): one-hot bool batches as the reference's generators yield them
    (helpers/ground_truth_conversion_utils.py:84-88, batch_generator_KITTI.py:82-84) -> one class id per pixel; any row
    that is not exactly one-hot makes it report 1 so that the feed ships the batch unchanged."""
    from fcn8s_tensorflow_b200 import ops
    rng = np.random.default_rng(3)
    for Cn, shape in ((20, (2, 64, 1024)), (2, (1, 300, 301)), (3, (1, 7, 9)), (7, (2, 5, 5)), (8, (1, 130, 130)),
                      (9, (1, 64, 1027)), (33, (1, 50, 50)), (255, (1, 9, 9)), (1, (1, 4, 4))):
        ids = rng.integers(0, Cn, size=shape).astype(np.uint8)
        onehot = np.eye(Cn, dtype=bool)[ids]
        out = np.full(shape, 251, np.uint8)
        for threads in (1, 4):
            assert ops.pack_labels(onehot, out, threads) and np.array_equal(out, ids), (Cn, threads)
        u8 = np.ascontiguousarray(onehot.view(np.uint8))
        assert ops.pack_labels(u8, out) and np.array_equal(out, ids)
        for where in ((0, 0, 0), (shape[0] - 1, shape[1] - 1, shape[2] - 1), (0, shape[1] // 2, shape[2] // 3)):
            empty = onehot.copy()
            empty[where] = False                      # a pixel without a class
            assert not ops.pack_labels(empty, out)
            two = u8.copy()
            two[where][int(ids[where])] = 2           # a label value other than 0 / 1
            assert not ops.pack_labels(two, out)
            if Cn > 1:
                multi = onehot.copy()
                multi[where][(int(ids[where]) + 1) % Cn] = True     # two classes on one pixel
                assert not ops.pack_labels(multi, out)
                moved = onehot.copy()                 # right number of set bytes, one of them in the neighbour's row
                flat = moved.reshape(-1, Cn)
                if flat.shape[0] > 1:
                    flat[0, :] = False
                    flat[1, :] = False
                    flat[1, 0] = True
                    if Cn > 1:
                        flat[1, Cn - 1] = True
                    assert not ops.pack_labels(moved, out)
    with pytest.raises(ValueError):
        ops.pack_labels(np.zeros((2, 2, 3), np.float32), np.zeros((2, 2), np.uint8))


def test_ctypes_structs_match_header_layout():
    """Field order / count of the ctypes mirrors against the C structs in the header."""
    from fcn8s_tensorflow_b200 import _capi
    hdr = open(os.path.join(ROOT, "include", "fcn8s_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for cname, cls in [("Fcn8PreprocessParams", _capi.PreprocessParams), ("Fcn8ConvParams", _capi.ConvParams),
                       ("Fcn8WgradParams", _capi.WgradParams), ("Fcn8PackParams", _capi.PackParams),
                       ("Fcn8PoolParams", _capi.PoolParams), ("Fcn8BiasGradParams", _capi.BiasGradParams),
                       ("Fcn8DeconvParams", _capi.DeconvParams), ("Fcn8Conv1Params", _capi.Conv1Params)]:
        body = dict((n, b) for b, n in re.findall(r"typedef struct \{([^}]*)\}\s*(\w+);", hdr))[cname]
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1]
            fields.append(first.lstrip("*"))
            fields += [n.strip().lstrip("*") for n in names[1:]]
        assert fields == [f[0] for f in cls._fields_], cname


def test_flat_layout_is_aligned_complete_and_backward_ordered():
    from fcn8s_tensorflow_b200.engine import flat_layout, variable_shapes
    layout, total = flat_layout(20)
    shapes = variable_shapes(20)
    assert set(layout) == set(shapes) and len(layout) == 42
    assert sum(int(np.prod(s)) for s in shapes.values()) == 134473144
    end = 0
    for name, (off, shape) in layout.items():
        assert off % 64 == 0 and off >= end
        end = off + int(np.prod(shape))
    assert total >= end
    names = list(layout)
    assert names[0].startswith("fc7_pool4_pool3_conv2d_trans") and names[-1] == "conv1_1/biases"
    assert names.index("fc7/weights") < names.index("fc6/weights") < names.index("conv5_3/filter")


def test_variable_shapes_agree_with_oracle():
    from fcn8s_tensorflow_b200.engine import variable_shapes
    from oracle import fcn8s_oracle as oracle
    for c in (2, 3, 20):
        assert dict(variable_shapes(c)) == dict(oracle.variable_shapes(c))


def test_dropout_rng_host_replica_statistics():
    from fcn8s_tensorflow_b200.rng import dropout_keep_mask, mix32
    m = dropout_keep_mask(1234, 1 << 16, 0.5)
    assert abs(m.float().mean().item() - 0.5) < 0.01
    assert dropout_keep_mask(1234, 64, 1.0).all()
    a = mix32(1, np.arange(8, dtype=np.uint64))
    b = mix32(2, np.arange(8, dtype=np.uint64))
    assert a.dtype == np.uint32 and (a != b).any()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_gpu():
    from fcn8s_tensorflow_b200 import _capi
    from fcn8s_tensorflow_b200.engine import Engine
    from fcn8s_tensorflow_b200.fcn8s import FCN8s
    with pytest.raises(_capi.Fcn8Error):
        Engine(20)
    with pytest.raises(_capi.Fcn8Error):
        FCN8s(vgg16_dir="synthetic", num_classes=2)
    with pytest.raises(ValueError):
        FCN8s()          # fcn8s_tensorflow.py:40-41


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fcn8s_tensorflow_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_hot_path_modules_call_no_library_compute():
    """The training / inference path (engine, ops, feed, dist, the class surface) must reach the GPU only through the
    C ABI: no torch.nn / functional / matmul / compile, no Triton, no cuDNN -- torch is the owner of device memory,
    streams and the process group, nothing else."""
    pkg = os.path.join(ROOT, "fcn8s_tensorflow_b200")
    banned = re.compile(r"torch\.nn\b|torch\.compile|import\s+triton|cudnn|\bF\.(conv|max_pool|softmax|cross_entropy|relu)"
                        r"|\.matmul\(|torch\.(mm|bmm|einsum|conv2d|conv_transpose2d|softmax|argmax)\(")
    for name in ("engine.py", "ops.py", "feed.py", "dist.py", "fcn8s.py", "_capi.py"):
        src = open(os.path.join(pkg, name)).read()
        hits = [m.group(0) for m in banned.finditer(src)]
        assert not hits, (name, hits)


def test_bench_line_carries_the_contract_keys():
    """bench.py's JSON line is assembled by train_line(): the driver's contract keys, computed from a fake result on
    the CPU (the timing itself needs a GPU)."""
    import json
    import types
    sys.path.insert(0, ROOT)
    import bench
    args = types.SimpleNamespace(steps=20, config="c2", backward_terms=3)
    flops = 2.0 * 4 * 890e9      # conv fprop + dgrad of 4 images, roughly
    r = dict(ms=290.0, launches=2020, e2e_ms=292.0, h2d=8388608, d2h=8, clocks={"sm_mhz": 1700.0}, loss=2.99, mem_gb=8.1,
             kernels={"conv_gemm": dict(launches=140, flops=5 * flops, ms=5 * 7.5),
                      "wgrad_gemm": dict(launches=80, flops=5 * flops / 2, ms=5 * 4.5), "_steps": 5})
    peaks = dict(tflops=1362.0, tflops_burst=1625.8, hbm=6548.8, source="test")
    line = bench.train_line(bench.CONFIGS["c2"], "fp32", r, args, 1, peaks)
    json.dumps(line)
    assert abs(line["value"] - 4 * 20 / 0.290) < 1e-6 and abs(line["ms_per_step"] - 14.5) < 1e-9
    assert line["gpu_launches"] == 2020 and line["e2e"]["h2d_bytes_per_step"] == 8388608
    ro = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in ro
    assert ro["bound"] == "tensor" and ro["unit"] == "TFLOP/s" and abs(ro["frac"] - ro["achieved"] / 1362.0) < 1e-12
    # the fp32-equivalent mode issues three MMAs per algorithmic product: its fraction is bounded by 1/3 of the pipe
    assert ro["mma_per_product"] == 3 and abs(ro["tensor_pipe_tflops"] - 3 * ro["achieved"]) < 1e-9
    assert bench.train_line(bench.CONFIGS["c2"], "bf16", r, args, 1, peaks)["roofline"]["mma_per_product"] == 1


def test_shard_bounds():
    from fcn8s_tensorflow_b200.dist import shard_batch, shard_bounds
    assert [shard_bounds(16, r, 8) for r in (0, 7)] == [(0, 2), (14, 16)]
    with pytest.raises(ValueError):
        shard_bounds(6, 0, 4)
    x = np.arange(8)
    parts = [shard_batch(x, x, r, 4)[0] for r in range(4)]
    assert np.array_equal(np.concatenate(parts), x)


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from fcn8s_tensorflow_b200 import dist as fdist
    r, _, w = fdist.init("gloo")

    class FakeEngine:
        pass
    e = FakeEngine()
    e.params = torch.full((10,), float(rank))
    e.adam_m = torch.zeros(10)
    e.adam_v = torch.zeros(10)
    e.world, e.allreduce, e._packed_dirty = 1, None, False
    e.global_step = 7 + rank          # restored from a checkpoint on rank 0 only: the broadcast makes it 7 everywhere
    fdist.attach(e)
    fdist.broadcast_parameters(e, src=0)
    assert e.global_step == 7 and e.rank == rank
    pair, conf = fdist.all_reduce_metrics(torch.tensor([1.5 * (rank + 1), 1.0], dtype=torch.float64),
                                          torch.full((2, 2), rank + 1, dtype=torch.int64))
    assert pair.tolist() == [4.5, 2.0] and conf.tolist() == [[3, 3], [3, 3]]
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    e.allreduce.start(g[:6])       # the engine's two-chunk protocol: head under the backward pass, tail before Adam
    e.allreduce.start(g[6:])
    e.allreduce.finish()
    assert not e.allreduce.pending
    avg = g / e.world     # what Adam's grad_scale = 1/world does
    t = fdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
    out.put((rank, e.world, e.params.tolist(), avg.tolist(), t, e._packed_dirty))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_host_logic_gloo_world2():
    """N>1 path on CPU: attach() + one all-reduce of the flat gradient + broadcast of the replicas, gloo backend."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_avg = (np.arange(10) * 1.5).tolist()     # mean of g*(1) and g*(2)
    for rank, world, params, avg, tmax, dirty in res:
        assert world == 2
        assert params == [0.0] * 10                  # broadcast from rank 0
        assert np.allclose(avg, expect_avg)
        assert tmax == 2.0 and dirty


# ---------------------------------------------------------------------------------------------------------------------
# Feed protocol against the reference's OWN generators (they import here; /root/reference does not exist on the GPU box)
REFERENCE = "/root/reference"


def _png_tree(root, n, h, w, classes, kitti=False):
    from PIL import Image
    rng = np.random.default_rng(0)
    os.makedirs(os.path.join(root, "images", "a"))
    os.makedirs(os.path.join(root, "labels", "a"))
    ids = []
    for i in range(n):
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        Image.fromarray(img).save(os.path.join(root, "images", "a", "um_%06d_leftImg8bit.png" % i))
        gt = rng.integers(0, classes, size=(h, w), dtype=np.uint8)
        ids.append(gt)
        if kitti:   # RGB labels named *_road_*, background = [255, 0, 0] (batch_generator_KITTI.py:39-44,82)
            rgb = np.zeros((h, w, 3), np.uint8)
            rgb[gt == 0] = (255, 0, 0)
            rgb[gt != 0] = (255, 0, 255)
            Image.fromarray(rgb).save(os.path.join(root, "labels", "a", "um_road_%06d_leftImg8bit.png" % i))
        else:
            Image.fromarray(gt).save(os.path.join(root, "labels", "a", "um_%06d_gtFine_labelIds.png" % i))
    return ids


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only mounted in the build container")
def test_reference_batch_generators_feed_the_class_surface(tmp_path):
    """The reference's BatchGenerator / KITTI batch_generator, run UNMODIFIED through the scipy.misc shim over a
    synthetic PNG tree, yield batches that the engine's feed checks accept as they are (uint8 images, bool one-hot
    labels, short last batch; batch_generator.py:244,390-391,414-415; batch_generator_KITTI.py:82-86,104-105)."""
    from fcn8s_tensorflow_b200.compat import install_scipy_misc_shim
    from fcn8s_tensorflow_b200.fcn8s import check_images, check_labels
    install_scipy_misc_shim()
    sys.path.insert(0, REFERENCE)
    try:
        from data_generator.batch_generator import BatchGenerator
        from data_generator.batch_generator_KITTI import batch_generator
    finally:
        sys.path.remove(REFERENCE)
    C, H, W = 5, 64, 96
    root = str(tmp_path / "cs")
    ids = _png_tree(root, 5, H, W, C)
    gen = BatchGenerator(image_dirs=[os.path.join(root, "images")], image_file_extension="png",
                         ground_truth_dirs=[os.path.join(root, "labels")], image_name_split_separator="leftImg8bit",
                         ground_truth_suffix="gtFine_labelIds", check_existence=True, num_classes=C)
    assert gen.get_num_files() == 5
    g = gen.generate(batch_size=2, convert_to_one_hot=True, shuffle=False)
    sizes = []
    for _ in range(4):                       # 2, 2, 1 (short last batch), then the next pass starts
        images, labels = next(g)
        sizes.append(len(images))
        assert images.dtype == np.uint8 and images.shape[1:] == (H, W, 3)
        assert labels.dtype == np.bool_ and labels.shape == (len(images), H, W, C)
        assert (labels.sum(-1) == 1).all()
        assert check_images(images) is images and check_labels(labels, C) is labels     # accepted without a copy
    assert sizes == [2, 2, 1, 2]
    # one-hot content equals the ids on disk (file order = sorted glob order is not guaranteed: compare as sets)
    g2 = gen.generate(batch_size=5, convert_to_one_hot=True, shuffle=False)
    _, labels = next(g2)
    got = sorted(l.argmax(-1).astype(np.uint8).tobytes() for l in labels)
    assert got == sorted(i.tobytes() for i in ids)
    with pytest.raises(ValueError):
        check_labels(labels[..., :3], C)
    with pytest.raises(ValueError):
        check_images(np.zeros((2, H, W), np.uint8))
    # KITTI road generator: 2 classes [background, road], resized to a x32 size
    kroot = str(tmp_path / "kitti")
    _png_tree(kroot, 3, 50, 70, 2, kitti=True)
    kg = batch_generator(2, kroot, "images/a", "labels/a", image_size=(64, 96))
    images, labels = next(kg)
    assert images.dtype == np.uint8 and images.shape == (2, 64, 96, 3)
    assert labels.dtype == np.bool_ and labels.shape == (2, 64, 96, 2) and (labels.sum(-1) == 1).all()
    assert check_labels(labels, 2) is labels


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only mounted in the build container")
def test_shipped_kitti_generator_equals_the_references_batch_for_batch(tmp_path):
    """fcn8s_tensorflow_b200.generators.batch_generator (what bench.py --config c5 feeds from on the GPU box, where
    /root/reference does not exist) against data_generator/batch_generator_KITTI.py:8-107 run through the scipy.misc
    shim: same files, same shuffles (python `random`), same flips (numpy RNG), bit-identical batches over two passes,
    including the short last batch and the bilinear label resize before the colour match."""
    import random
    from fcn8s_tensorflow_b200.compat import install_scipy_misc_shim
    from fcn8s_tensorflow_b200 import generators
    install_scipy_misc_shim()
    sys.path.insert(0, REFERENCE)
    try:
        from data_generator.batch_generator_KITTI import batch_generator as ref_generator
    finally:
        sys.path.remove(REFERENCE)
    root = str(tmp_path / "kitti")
    idir, ldir = generators.write_synthetic_kitti_tree(root, 5, height=75, width=142, seed=3)
    for flip in (False, 0.5):
        random.seed(11)
        np.random.seed(5)
        ref = ref_generator(2, root, idir, ldir, image_size=(64, 128), flip=flip)
        want = [next(ref) for _ in range(7)]
        random.seed(11)
        np.random.seed(5)
        mine = generators.batch_generator(2, root, idir, ldir, image_size=(64, 128), flip=flip)
        got = [next(mine) for _ in range(7)]
        assert [len(b[0]) for b in got] == [2, 2, 1, 2, 2, 1, 2]
        for (wi, wl), (gi, gl) in zip(want, got):
            assert gi.dtype == np.uint8 and gl.dtype == np.bool_ and gl.shape[-1] == 2
            assert np.array_equal(wi, gi) and np.array_equal(wl, gl)
        assert any(gl[..., 1].any() for _, gl in got) and any(gl[..., 0].any() for _, gl in got)
    # images only (labels_subdir=None) yields bare image batches, like the reference (:106-107)
    random.seed(1)
    only = next(generators.batch_generator(3, root, idir, None, image_size=(32, 64)))
    assert isinstance(only, np.ndarray) and only.shape == (3, 32, 64, 3)


def test_variable_summaries_match_tensorflow_buckets_and_reference_tags(tmp_path):
    """helpers/tf_variable_summaries.py:3-20 + fcn8s_tensorflow.py:331-350: mean / population stddev / max / min and a
    histogram over TensorFlow's default buckets, one set per variable under the reference's scopes."""
    import torch
    from fcn8s_tensorflow_b200 import summaries as S
    limits = S.tf_bucket_limits()
    assert len(limits) == 1550 and limits[774] == 0.0 and limits[775] == 1e-12 and limits[773] == -1e-12
    assert np.all(np.diff(limits) > 0) and abs(limits[776] / limits[775] - 1.1) < 1e-12
    g = torch.Generator().manual_seed(0)
    x = torch.randn(7, 5, 300, generator=g) * 0.01
    x[0, 0, 0] = 0.0
    x[0, 0, 1] = 1e-12                      # exactly on a bucket limit: belongs to the bucket that STARTS there
    st = S.variable_stats(x).numpy()
    xn = x.numpy().astype(np.float64)
    assert np.allclose(st, [xn.mean(), np.sqrt(np.mean((xn - xn.mean()) ** 2)), xn.max(), xn.min()], rtol=1e-6)
    h = S.variable_histogram(x)
    counts, _ = np.histogram(xn.reshape(-1), bins=limits)       # [edge_j, edge_j+1): bucket j+1 of the TF scheme
    full = np.concatenate([[0], counts])
    nz = np.nonzero(full)[0]
    assert h["bucket_limits"] == limits[nz[0]:nz[-1] + 1].tolist()
    assert h["bucket_counts"] == full[nz[0]:nz[-1] + 1].astype(float).tolist()
    assert h["num"] == xn.size and abs(h["sum"] - xn.sum()) < 1e-9 and abs(h["sum_squares"] - (xn ** 2).sum()) < 1e-9
    assert h["min"] == xn.min() and h["max"] == xn.max()

    from torch.utils.tensorboard import SummaryWriter
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    names = [n for n, _ in S.SUMMARY_VARIABLES]
    assert len(names) == 20 and names[12:16] == ["fc7/weights", "fc7/biases", "fc6/weights", "fc6/biases"]
    tensors = {n: torch.randn(4, 3, generator=g) for n in names}
    w = SummaryWriter(str(tmp_path))
    S.write_variable_summaries(w, tensors, 11)
    w.close()
    acc = EventAccumulator(str(tmp_path), size_guidance={"histograms": 0, "scalars": 0})
    acc.Reload()
    tags = acc.Tags()
    for _, scope in S.SUMMARY_VARIABLES:
        for t in ("mean", "stddev", "max", "min"):
            assert "%s/%s" % (scope, t) in tags["scalars"]
        assert "%s/histogram" % scope in tags["histograms"]
    ev = acc.Scalars("fc6/kernel/max")[0]
    assert ev.step == 11 and abs(ev.value - float(tensors["fc6/weights"].max())) < 1e-6
    hv = acc.Histograms("conv3_3/bias/histogram")[0].histogram_value
    assert hv.num == 12 and abs(hv.sum - float(tensors["conv3_3/biases"].double().sum())) < 1e-6


def test_overlay_matches_the_references_helper_bit_for_bit():
    """`print_segmentation_onto_image` (helpers/visualization_utils.py:7-52) through the fixture generated by importing
    the reference's own function (tests/golden/make_golden.py::reference_overlay)."""
    from fcn8s_tensorflow_b200.fcn8s import print_segmentation_onto_image
    g = np.load(os.path.join(ROOT, "tests", "golden", "reference_overlay.npz"))
    color_map = {int(k): tuple(int(x) for x in v) for k, v in zip(g["classes"], g["colors"])}
    out = print_segmentation_onto_image(g["image"], g["segmentation"], color_map)
    assert out.dtype == np.uint8 and np.array_equal(out, g["overlay"])
    untouched = g["segmentation"] == 0                     # class 0 is not in the colour map
    assert np.array_equal(out[untouched], g["image"][untouched])
    with pytest.raises(ValueError, match="must have the same height and width"):
        print_segmentation_onto_image(g["image"], g["segmentation"][:-1], color_map)


def test_predict_and_save_file_handling_without_a_device(tmp_path):
    """fcn8s_tensorflow.py:772-855 with `predict` stubbed: every `*.png` of images_dir is read, optionally resized,
    overlaid, optionally stacked with the unprocessed image (vertical / horizontal) and written under the same name;
    an existing results directory is replaced."""
    from PIL import Image
    from fcn8s_tensorflow_b200.fcn8s import FCN8s, print_segmentation_onto_image
    rng = np.random.default_rng(3)
    src = tmp_path / "images"
    src.mkdir()
    imgs = {}
    for name, (h, w) in {"a.png": (32, 64), "b.png": (48, 32)}.items():
        imgs[name] = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        Image.fromarray(imgs[name]).save(str(src / name))
    (src / "ignored.jpg").write_bytes(b"not an image")
    color_map = {1: (0, 255, 0, 127), 2: (255, 0, 0, 200)}

    def fake_predict(images, argmax=True):
        assert argmax is True and len(images) == 1
        return np.stack([(im[..., 0] // 86).astype(np.int64) for im in images])      # classes 0..2 from the red channel

    m = object.__new__(FCN8s)          # no engine, no device: only the file handling is under test
    m.predict = fake_predict
    out = tmp_path / "results"
    out.mkdir()
    (out / "stale.txt").write_text("x")
    with contextlib.redirect_stdout(io.StringIO()):
        m.predict_and_save(str(out), str(src), color_map)
    assert sorted(os.listdir(out)) == ["a.png", "b.png"]
    for name, im in imgs.items():
        got = np.asarray(Image.open(str(out / name)).convert("RGB"))
        assert np.array_equal(got, print_segmentation_onto_image(im, fake_predict([im])[0], color_map))
    with contextlib.redirect_stdout(io.StringIO()):
        m.predict_and_save(str(out), str(src), color_map, include_unprocessed_image=True, arrangement='vertical')
    got = np.asarray(Image.open(str(out / "a.png")).convert("RGB"))
    assert got.shape == (64, 64, 3) and np.array_equal(got[32:], imgs["a.png"])
    with contextlib.redirect_stdout(io.StringIO()):
        m.predict_and_save(str(out), str(src), color_map, resize=(32, 32), include_unprocessed_image=True,
                           arrangement='horizontal')
    got = np.asarray(Image.open(str(out / "b.png")).convert("RGB"))
    resized = np.asarray(Image.open(str(src / "b.png")).convert("RGB").resize((32, 32), Image.BILINEAR))
    assert got.shape == (32, 64, 3) and np.array_equal(got[:, 32:], resized)
    assert np.array_equal(got[:, :32], print_segmentation_onto_image(resized, fake_predict([resized])[0], color_map))
