"""KITTI-road batch generator with the reference's protocol (data_generator/batch_generator_KITTI.py:8-107).

Same signature, same yield contract -- `(images uint8 [n,H,W,3], labels bool [n,H,W,2])`, channel 0 = background
(label colour [255,0,0], :44,82), channel 1 = everything else; images AND labels are resized to `image_size` with the
bilinear filter `scipy.misc.imresize` used (:71,78; labels are interpolated before the colour match, a reference quirk
kept on purpose); optional horizontal flip with probability `flip` (:96-101); short last batch of
a pass; reshuffle after every complete pass (:56-58).

Why it exists: the reference's generator imports `scipy.misc.imread / imresize`, which SciPy removed, and
/root/reference is not present on the GPU box, so `bench.py --config c5` (BASELINE configs[4], "batch_generator_KITTI
feed path") needs a generator that ships with the product.  `tests/test_host_cpu.py` checks it batch-for-batch against
the reference's own generator (run through `compat.install_scipy_misc_shim`) where the reference tree is mounted.
Decoding + resizing of the files runs on a small thread pool (PIL releases the GIL), one batch AHEAD: the files of
batch k+1 are being decoded while batch k is assembled, yielded and consumed, so the consumer sees the decode latency
of a batch only once.  The batches, their order, the reshuffle points and the use of the random generators are the
reference's (the path list of batch k+1 is taken -- and, at the end of a pass, reshuffled -- before batch k is yielded
instead of after; nothing else in the generator draws from `random`).
"""
import os
import random
import re
from concurrent.futures import ThreadPoolExecutor
from glob import glob

import numpy as np

BACKGROUND_COLOR = np.array([255, 0, 0])     # batch_generator_KITTI.py:44
_POOL = ThreadPoolExecutor(max_workers=8, thread_name_prefix="fcn8-kitti-decode")


def _load_resized(path, image_size):
    """scipy.misc.imresize(scipy.misc.imread(path), image_size): PIL open, bilinear resize to (height, width)."""
    from PIL import Image
    img = Image.open(path)
    if img.mode not in ("RGB", "L", "RGBA"):
        img = img.convert("RGB")
    a = np.array(img.resize((int(image_size[1]), int(image_size[0])), resample=Image.BILINEAR))
    return a


def _load_label(path, image_size):
    """Resized label image -> bool [H, W, 2]: channel 0 = pixels of exactly the background colour, channel 1 = the rest
    (batch_generator_KITTI.py:78-84).  Runs on the decode pool like the image files."""
    label = _load_resized(path, image_size)
    if label.ndim == 3 and label.shape[2] >= 3:
        background = ((label[..., 0] == BACKGROUND_COLOR[0]) & (label[..., 1] == BACKGROUND_COLOR[1]) &
                      (label[..., 2] == BACKGROUND_COLOR[2]))
        if label.shape[2] > 3:     # RGBA file: np.all(label == [255, 0, 0], axis=2) cannot broadcast in the reference
            raise ValueError("label image %s has %d channels" % (path, label.shape[2]))
        background = background[..., None]
    else:
        background = np.all(label == BACKGROUND_COLOR, axis=2)[..., None]
    return np.concatenate((background, np.invert(background)), axis=2)


def batch_generator(batch_size, dataset_rootdir, images_subdir, labels_subdir, image_size, flip=False):
    image_paths = glob(os.path.join(dataset_rootdir, images_subdir, '*.png'))
    label_paths = None
    if labels_subdir is not None:
        label_paths = {re.sub(r'_road_', '_', os.path.basename(path)): path
                       for path in glob(os.path.join(dataset_rootdir, labels_subdir, '*_road_*.png'))}
    random.shuffle(image_paths)
    state = {"current": 0}

    def start_next_batch():
        """The reference's loop head (:56-61) for the next batch + its decode jobs."""
        if state["current"] >= len(image_paths):
            random.shuffle(image_paths)
            state["current"] = 0
        paths = image_paths[state["current"]:state["current"] + batch_size]
        state["current"] += batch_size
        jobs = [_POOL.submit(_load_resized, p, image_size) for p in paths]
        if label_paths is not None:
            jobs += [_POOL.submit(_load_label, label_paths[os.path.basename(p)], image_size) for p in paths]
        return paths, jobs

    ahead = start_next_batch()
    while True:
        paths, jobs = ahead
        ahead = start_next_batch()
        done = [j.result() for j in jobs]
        images = done[:len(paths)]
        labels = done[len(paths):]
        for i in range(len(images)):
            if flip:
                p = np.random.uniform(0, 1)
                if p >= (1 - flip):
                    images[i] = images[i][:, ::-1, :]
                    if label_paths is not None:
                        labels[i] = labels[i][:, ::-1, :]
        if label_paths is not None:
            yield np.array(images), np.array(labels)
        else:
            yield np.array(images)


def write_synthetic_kitti_tree(root, count, height=375, width=1242, seed=0):
    """A KITTI-road-shaped PNG tree for benchmarks / tests (no dataset is available offline): `training/image_2/
    um_%06d.png` RGB images and `training/gt_image_2/um_road_%06d.png` labels whose background is [255,0,0] and whose
    road region (a random trapezoid) is [255,0,255], the dataset's colours.  Returns (images_subdir, labels_subdir)."""
    from PIL import Image, ImageDraw
    rng = np.random.default_rng(seed)
    idir, ldir = os.path.join("training", "image_2"), os.path.join("training", "gt_image_2")
    os.makedirs(os.path.join(root, idir), exist_ok=True)
    os.makedirs(os.path.join(root, ldir), exist_ok=True)
    for i in range(count):
        # smooth-ish content so that the PNGs compress like photographs rather than like noise
        base = rng.integers(0, 256, size=(height // 8 + 1, width // 8 + 1, 3), dtype=np.uint8)
        img = np.asarray(Image.fromarray(base).resize((width, height), resample=Image.BILINEAR))
        img = np.clip(img.astype(np.int16) + rng.integers(-6, 7, size=img.shape, dtype=np.int16), 0, 255).astype(np.uint8)
        Image.fromarray(img).save(os.path.join(root, idir, "um_%06d.png" % i))
        lab = Image.new("RGB", (width, height), (255, 0, 0))
        x0, x1 = sorted(rng.integers(width // 4, 3 * width // 4, size=2).tolist())
        ImageDraw.Draw(lab).polygon([(x0, height // 2), (x1 + 40, height // 2), (width - 1, height - 1), (0, height - 1)],
                                    fill=(255, 0, 255))
        lab.save(os.path.join(root, ldir, "um_road_%06d.png" % i))
    return idir, ldir
