"""Compatibility shim that lets the reference's data generators run unmodified on current SciPy.

`data_generator/batch_generator.py:247,254,399,406` and `data_generator/batch_generator_KITTI.py:71,78` call
`scipy.misc.imread / imresize / imsave`, and `helpers/visualization_utils.py:45-47` `scipy.misc.toimage`, which SciPy
removed in 1.2/1.3.  `install_scipy_misc_shim()` adds PIL-backed
stand-ins with the same call signatures to the `scipy.misc` module (only the names that are missing), so a user keeps
feeding `FCN8s.train / evaluate` from `BatchGenerator(...).generate(...)` / `batch_generator(...)` as before.
"""
import numpy as np


def _imread(path, flatten=False, mode=None):
    from PIL import Image
    img = Image.open(path)
    if mode is not None:
        img = img.convert(mode)
    elif flatten:
        img = img.convert('F')
    return np.array(img)


def _imresize(arr, size, interp='bilinear', mode=None):
    """scipy.misc.imresize semantics: `size` is (height, width), a fraction or a percentage; returns uint8."""
    from PIL import Image
    a = np.asarray(arr)
    img = Image.fromarray(a if a.dtype == np.uint8 else _bytescale(a), mode=mode)
    if isinstance(size, (int, np.integer)):
        frac = size / 100.0
        new = (int(img.width * frac), int(img.height * frac))
    elif isinstance(size, float):
        new = (int(img.width * size), int(img.height * size))
    else:
        new = (int(size[1]), int(size[0]))
    method = {'nearest': Image.NEAREST, 'lanczos': Image.LANCZOS, 'bilinear': Image.BILINEAR,
              'bicubic': Image.BICUBIC, 'cubic': Image.BICUBIC}[interp]
    return np.array(img.resize(new, resample=method))


def _bytescale(a):
    a = a.astype(np.float64)
    lo, hi = a.min(), a.max()
    if hi == lo:
        return np.zeros(a.shape, np.uint8)
    return ((a - lo) * (255.0 / (hi - lo)) + 0.5).clip(0, 255).astype(np.uint8)


def _imsave(path, arr, format=None):
    from PIL import Image
    a = np.asarray(arr)
    Image.fromarray(a if a.dtype == np.uint8 else _bytescale(a)).save(path, format=format)


def _toimage(arr, high=255, low=0, cmin=None, cmax=None, pal=None, mode=None, channel_axis=None):
    """scipy.misc.toimage for the cases the reference uses: uint8 [H,W], [H,W,3] or [H,W,4] arrays become PIL images
    unchanged (SciPy's byte-scaling is the identity on uint8 data); other dtypes are byte-scaled first."""
    from PIL import Image
    a = np.asarray(arr)
    if a.dtype != np.uint8:
        a = _bytescale(a)
    if mode is None:
        mode = {2: 'L', 3: {3: 'RGB', 4: 'RGBA'}.get(a.shape[-1])}.get(a.ndim)
    return Image.fromarray(a, mode=mode)


def install_scipy_misc_shim():
    """Add imread / imresize / imsave / toimage to scipy.misc when they are missing. Returns the names added."""
    import scipy.misc as misc
    added = []
    for name, fn in (("imread", _imread), ("imresize", _imresize), ("imsave", _imsave), ("toimage", _toimage)):
        if not hasattr(misc, name):
            setattr(misc, name, fn)
            added.append(name)
    return added
