"""`FCN8s` -- the reference's model/trainer class surface (fcn8s_tensorflow.py:17-952) on the B200 engine.

Same constructor, `train` / `evaluate` / `predict` / `predict_and_save` / `save` / `load_variables` / `close`
signatures, the same public attributes (`metric_names`, `metric_values`, `best_metric_values`, `training_loss`,
`best_training_loss`, `g_step`, `variables_updated`, `eval_dataset`), the same argument validation and error messages,
and the same generator protocol (`next(gen)` -> uint8 images [n,H,W,3], bool one-hot labels [n,H,W,C];
data_generator/batch_generator.py:414-415).  What replaces `tf.Session.run` is `engine.Engine`.

Differences that are inherent to leaving TensorFlow (documented in INTEGRATION.md):
  * `vgg16_dir` / `model_load_dir` / `variables_load_dir` are read as TensorFlow tensor bundles (a SavedModel's
    `variables/variables.index` + `.data-00000-of-00001`, or a `train_saver` prefix) by `tf_bundle.py`, without
    TensorFlow and by TF variable name; `.npz` files keyed the same way and the `synthetic[:seed]` scheme also work.
    The TF1 graph in `saved_model.pb` is never read: the graph is the engine.
  * `save()` writes the same variable files under the reference's directory names (`variables/variables.*` for
    `saver='saved_model'`, `variables.*` + `checkpoint` for `'train_saver'`), including the Adam slots and
    `optimizer/global_step`; it does not write a `saved_model.pb`.
  * extra keyword-only constructor arguments select the precision mode and the device.
"""
import os
import shutil
import sys
import time
import warnings
from collections import deque
from glob import glob

import numpy as np
import torch

from . import _capi as capi
from .engine import Engine, variable_shapes
from .summaries import SUMMARY_VARIABLES, write_variable_summaries


def synthetic_weights(num_classes, seed=2, decoder_std_scale=1.0):
    """Deterministic stand-in for the pretrained VGG-16 SavedModel (README.md:42, not available offline): He-normal
    encoder, the reference's decoder initialisers (truncated normal sigma 1e-3 / 1e-2, zero biases;
    fcn8s_tensorflow.py:159-160,178,209)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape in variable_shapes(num_classes).items():
        decoder = not (name.startswith("conv") or name.startswith("fc6") or name.startswith("fc7/"))
        if len(shape) == 1:
            v = np.zeros(shape, np.float32) if decoder else (rng.standard_normal(shape) * 0.01).astype(np.float32)
        elif decoder:
            std = (1e-3 if shape[0] == 1 else 1e-2) * decoder_std_scale
            x = rng.standard_normal(shape)
            bad = np.abs(x) > 2.0
            while bad.any():
                x[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(x) > 2.0
            v = (x * std).astype(np.float32)
        else:
            v = (rng.standard_normal(shape) * np.sqrt(2.0 / (shape[0] * shape[1] * shape[2]))).astype(np.float32)
        out[name] = v
    return out


def check_images(images):
    """The `image_input` feed (fcn8s_tensorflow.py:558,686,765): uint8-valued RGB [n,H,W,3]; a list of [H,W,3] arrays is
    accepted like `sess.run` accepts it.  uint8 arrays pass through without a copy."""
    a = np.asarray(images)
    if a.ndim != 4 or a.shape[-1] != 3:
        raise ValueError("images must have shape (batch, height, width, 3), got %s" % (a.shape,))
    if a.dtype != np.uint8:
        q = np.clip(np.rint(a), 0, 255)
        if not np.array_equal(q, a):
            warnings.warn("image feed is not uint8-valued; it is rounded / clipped to 0..255 (the reference feeds raw "
                          "8-bit images, fcn8s_tensorflow.py:558)", stacklevel=3)
        a = q.astype(np.uint8)
    return a


def check_labels(labels, num_classes):
    """The `labels_input` feed (:110,559): one-hot [n,H,W,C], bool as the generators yield it
    (helpers/ground_truth_conversion_utils.py:84-88, batch_generator_KITTI.py:82-84) or any 0/1 integer type."""
    a = np.asarray(labels)
    if a.ndim != 4 or a.shape[-1] != num_classes:
        raise ValueError("labels must be one-hot with shape (batch, height, width, %d), got %s"
                         % (num_classes, a.shape,))
    if a.dtype != np.bool_ and a.dtype != np.uint8:
        a = a.astype(np.uint8)
    return a


def print_segmentation_onto_image(image, segmentation, color_map):
    """The overlay of helpers/visualization_utils.py:7-52 on a class-id map: an RGBA layer holding `color_map[class]`
    at every pixel of that class is pasted over the RGB image with itself as the mask (PIL `Image.paste`, the operation
    and therefore the rounding the reference uses).  `segmentation`: int [H, W] (the reference takes one-hot / softmax
    [1, H, W, C] and arg-maxes it first).  Returns uint8 [H, W, 3]."""
    from PIL import Image
    image = np.asarray(image, dtype=np.uint8)
    segmentation = np.asarray(segmentation)
    if image.shape[:2] != segmentation.shape[-2:]:
        raise ValueError("'image' and 'prediction' must have the same height and width, but image has spatial "
                         "dimensions ({}, {}) and prediction has spatial dimensions ({}, {}).".format(
                             image.shape[0], image.shape[1], segmentation.shape[-2], segmentation.shape[-1]))
    layer = np.zeros((image.shape[0], image.shape[1], 4), dtype=np.uint8)
    for segmentation_class, color_value in color_map.items():
        layer[segmentation == segmentation_class] = color_value
    layer = Image.fromarray(layer, mode="RGBA")
    out = Image.fromarray(image)
    out.paste(layer, box=None, mask=layer)
    return np.asarray(out, dtype=np.uint8)


def _load_npz_weights(path, num_classes=None):
    """Variables by TF name from `path`: a TensorFlow tensor bundle -- a SavedModel directory
    (`variables/variables.index` + `.data-00000-of-00001`, what fcn8s_tensorflow.py:74,134 load and :922-925 writes), a
    `train_saver` directory or prefix (:926-944) -- read without TensorFlow by tf_bundle.py; or an .npz keyed the same
    way.  The TF1 graph in `saved_model.pb` is not needed: the graph is this engine."""
    from . import tf_bundle
    prefix = tf_bundle.find_bundle(path)
    if prefix is not None:
        return dict(tf_bundle.read_bundle(prefix))
    if os.path.isdir(path):
        cands = [os.path.join(path, f) for f in ("variables.npz", "vgg16_weights.npz", "weights.npz")]
        found = [c for c in cands if os.path.exists(c)]
        if not found:
            raise FileNotFoundError("no TensorFlow variables (variables/variables.index) and no variables.npz / "
                                    "vgg16_weights.npz under %s" % path)
        path = found[0]
    data = np.load(path)
    return {k: data[k] for k in data.files}


class FCN8s:

    def __init__(self, model_load_dir=None, tags=None, vgg16_dir=None, num_classes=None, variables_load_dir=None, *,
                 precision="fp32", device=None, seed=2, weights=None, data_parallel=False, backward_terms=3,
                 grad_comm=None):
        # precision: "fp32" (default, like the reference: the fp32-equivalent mode that meets the 1e-4 logit contract)
        # or the opt-in reduced-precision "bf16" (BASELINE configs[2]); see engine.Engine.
        # fcn8s_tensorflow.py:40-41
        if (weights is None) and (model_load_dir is None) and (vgg16_dir is None or num_classes is None):
            raise ValueError("You must provide either both `model_load_dir` and `tags` or both `vgg16_dir` and `num_classes`.")

        self.variables_load_dir = variables_load_dir
        self.model_load_dir = model_load_dir
        self.tags = tags
        self.vgg16_dir = vgg16_dir
        self.vgg16_tag = 'vgg16'
        self.num_classes = num_classes

        self.variables_updated = False
        self.eval_dataset = None
        self.metric_names = []
        self.metric_values = []
        self.best_metric_values = []
        self.training_loss = None
        self.best_training_loss = 99999999.9
        self.g_step = None

        adam_state = None
        if weights is not None:
            self.num_classes = num_classes = int(np.asarray(weights["fc7_1x1/bias"]).shape[0])
        elif model_load_dir is not None:
            weights = _load_npz_weights(model_load_dir, None)
            self.num_classes = num_classes = int(weights["fc7_1x1/bias"].shape[0])
            adam_state = weights
        elif isinstance(vgg16_dir, str) and vgg16_dir.startswith("synthetic"):
            s = int(vgg16_dir.split(":")[1]) if ":" in vgg16_dir else seed
            weights = synthetic_weights(num_classes, s)
        else:
            weights = synthetic_weights(num_classes, seed)       # decoder initialisers (built from scratch, :108)
            enc = _load_npz_weights(vgg16_dir, num_classes)      # pretrained encoder (:106)
            for k in list(weights):
                if k.startswith("conv") or k.startswith("fc6") or k.startswith("fc7/"):
                    weights[k] = enc[k]

        self.engine = Engine(num_classes, precision=precision, device=device, backward_terms=backward_terms,
                             grad_comm=grad_comm)
        self.engine.load_weights(weights)
        if data_parallel:
            from . import dist as _dist
            _dist.attach(self.engine)
            _dist.broadcast_parameters(self.engine)
        if adam_state is not None:
            self._restore_optimizer(adam_state)
        if variables_load_dir is not None and model_load_dir is None:
            self.load_variables(variables_load_dir)
        if data_parallel:    # the optimiser state / global step restored above must also be identical on every rank
            _dist.broadcast_parameters(self.engine)
        self._conf = torch.zeros((num_classes, num_classes), dtype=torch.int64, device=self.engine.device)
        self._metric_set = set()
        self._stage = {}

    # ------------------------------------------------------------------ feed (fcn8s_tensorflow.py:558-562, 686-689, 765)
    def _to_device(self, array, key):
        """Host batch -> device through a pinned staging buffer (one per (key, shape)); async on the current stream."""
        a = np.ascontiguousarray(np.asarray(array))
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        k = (key, a.shape, a.dtype.str)
        st = self._stage.get(k)
        if st is None:
            pin = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype).pin_memory()
            dev = torch.empty(a.shape, dtype=pin.dtype, device=self.engine.device)
            st = self._stage[k] = (pin, dev)
        pin, dev = st
        pin.copy_(torch.from_numpy(a))
        dev.copy_(pin, non_blocking=True)
        return dev

    def _check_images(self, images):
        return check_images(images)

    def _check_labels(self, labels):
        return check_labels(labels, self.num_classes)

    def _images_to_device(self, images):
        return self._to_device(self._check_images(images), "images")

    def _labels_to_device(self, labels):
        return self._to_device(self._check_labels(labels), "labels")

    def _loss_read_begin(self, shape):
        """Async D2H of the step's loss pair into a pinned slot; returns a token for `_loss_read_end`."""
        if not hasattr(self, "_loss_slots"):
            self._loss_slots = [(torch.empty(2, dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(2)]
            self._loss_i = 0
        host, ev = self._loss_slots[self._loss_i]
        self._loss_i ^= 1
        host.copy_(self.engine.loss_buf, non_blocking=True)
        ev.record(torch.cuda.current_stream(self.engine.device))
        return host, ev, float(shape[0] * shape[1] * shape[2])

    @staticmethod
    def _loss_read_end(token):
        host, ev, npx = token
        ev.synchronize()
        return float(host[0]) / npx + float(host[1])

    # ------------------------------------------------------------------ metrics (fcn8s_tensorflow.py:273-322, 371-397)
    def _initialize_metrics(self, metrics):
        self.metric_names = []
        self.best_metric_values = []
        if 'loss' in metrics:
            self.metric_names.append('loss')
            self.best_metric_values.append(99999999.9)
        if 'mean_iou' in metrics:
            self.metric_names.append('mean_iou')
            self.best_metric_values.append(0.0)
        if 'accuracy' in metrics:
            self.metric_names.append('accuracy')
            self.best_metric_values.append(0.0)

    def _metric_values_from(self, loss_sum, loss_count, cm):
        vals = []
        for name in self.metric_names:
            if name == 'loss':
                vals.append(loss_sum / loss_count if loss_count else 0.0)
            elif name == 'mean_iou':
                cmf = cm.astype(np.float64)
                diag = np.diag(cmf)
                denom = cmf.sum(0) + cmf.sum(1) - diag
                valid = denom > 0
                vals.append(float((diag[valid] / denom[valid]).sum() / valid.sum()) if valid.any() else 0.0)
            elif name == 'accuracy':
                tot = cm.sum()
                vals.append(float(np.diag(cm).sum() / tot) if tot else 0.0)
        return vals

    # ------------------------------------------------------------------ train (fcn8s_tensorflow.py:399-658)
    def train(self,
              train_generator,
              epochs,
              steps_per_epoch,
              learning_rate_schedule,
              keep_prob=0.5,
              l2_regularization=0.0,
              eval_dataset='train',
              eval_frequency=5,
              val_generator=None,
              val_steps=None,
              metrics={},
              save_during_training=False,
              save_dir=None,
              save_best_only=True,
              save_tags=['default'],
              save_name='',
              save_frequency=5,
              saver='saved_model',
              monitor='loss',
              record_summaries=True,
              summaries_frequency=10,
              summaries_dir=None,
              summaries_name=None,
              training_loss_display_averaging=3):
        from tqdm import trange

        if not eval_dataset in ['train', 'val']:
            raise ValueError("`eval_dataset` must be one of 'train' or 'val', but is '{}'.".format(eval_dataset))
        if (eval_dataset == 'val') and ((val_generator is None) or (val_steps is None)):
            raise ValueError("When eval_dataset == 'val', a `val_generator` and `val_steps` must be passed.")
        for metric in metrics:
            if not metric in ['loss', 'mean_iou', 'accuracy']:
                raise ValueError("{} is not a valid metric. Valid metrics are ['loss', mean_iou', 'accuracy']".format(metric))
        if (not monitor in metrics) and (not monitor == 'loss'):
            raise ValueError('You are trying to monitor {}, but it is not in `metrics` and is therefore not being computed.'.format(monitor))

        self.eval_dataset = eval_dataset
        self.g_step = self.engine.global_step
        learning_rate = learning_rate_schedule(self.g_step)
        self._initialize_metrics(metrics)

        training_writer = evaluation_writer = None
        if record_summaries and summaries_dir is not None:
            # the reference crashes on summaries_dir=None (os.path.join(None, ...)); here that
            # combination simply records nothing
            from torch.utils.tensorboard import SummaryWriter
            training_writer = SummaryWriter(os.path.join(summaries_dir, summaries_name or 'summaries'))
            if len(metrics) > 0:
                evaluation_writer = SummaryWriter(os.path.join(summaries_dir, (summaries_name or 'summaries') + '_eval'))

        from .feed import Feeder
        for epoch in range(1, epochs + 1):
            loss_history = deque(maxlen=training_loss_display_averaging)
            tr = trange(steps_per_epoch, file=sys.stdout)
            tr.set_description('Epoch {}/{}'.format(epoch, epochs))
            # The feed is prefetched (feed.py) and each step's loss is fetched one step late, so the device never
            # waits for the host: `pending` is the (loss token, step, learning rate) of the previous step.
            feeder = Feeder(self, train_generator, steps_per_epoch)
            pending = None

            def account(p):
                current_loss = self._loss_read_end(p[0])
                if training_writer is not None and (p[1] - 1) % summaries_frequency == 0:
                    training_writer.add_scalar('total_loss', current_loss, p[1])
                    training_writer.add_scalar('learning_rate', p[2], p[1])
                loss_history.append(current_loss)
                self.training_loss = float(np.mean(np.array(loss_history)))
                tr.set_postfix(ordered_dict={'loss': self.training_loss, 'learning rate': p[2]})

            try:
                for train_step in tr:
                    x, y = feeder.get()
                    self.engine.train_step(x, y, learning_rate, keep_prob, l2_regularization)
                    feeder.release()
                    token = self._loss_read_begin(x.shape)
                    self.g_step = self.engine.global_step
                    self.variables_updated = True
                    if training_writer is not None and (self.g_step - 1) % summaries_frequency == 0:
                        # variable summaries of _build_summary_ops (:331-350), reduced on the device (summaries.py)
                        e = self.engine
                        write_variable_summaries(training_writer,
                                                 {n: e.view(n, e.params) for n, _ in SUMMARY_VARIABLES}, self.g_step)
                    if pending is not None:
                        account(pending)
                    pending = (token, self.g_step, learning_rate)
                    learning_rate = learning_rate_schedule(self.g_step)
                if pending is not None:
                    account(pending)
            finally:
                feeder.close()     # stops and joins the prefetch thread also when the loop raises

            if (len(metrics) > 0) and (epoch % eval_frequency == 0):
                if eval_dataset == 'train':
                    data_generator, num_batches, description = train_generator, steps_per_epoch, 'Evaluation on training dataset'
                else:
                    data_generator, num_batches, description = val_generator, val_steps, 'Evaluation on validation dataset'
                self._evaluate(data_generator=data_generator, metrics=metrics, num_batches=num_batches,
                               l2_regularization=l2_regularization, description=description)
                if evaluation_writer is not None:
                    for n_, v_ in zip(self.metric_names, self.metric_values):
                        evaluation_writer.add_scalar({'loss': 'mean_loss'}.get(n_, n_), v_, self.g_step)

            if save_during_training and (epoch % save_frequency == 0):
                save = False
                if save_best_only:
                    if (monitor == 'loss' and (not 'loss' in self.metric_names) and
                            self.training_loss < self.best_training_loss):
                        save = True
                    else:
                        i = self.metric_names.index(monitor)
                        if (monitor == 'loss') and (self.metric_values[i] < self.best_metric_values[i]):
                            save = True
                        # the reference spells this 'accuracry' (fcn8s_tensorflow.py:626), so monitoring accuracy
                        # never saves there; here accuracy is honoured.
                        elif (monitor in ['accuracy', 'mean_iou']) and (self.metric_values[i] > self.best_metric_values[i]):
                            save = True
                    if save:
                        print('New best {} value, saving model.'.format(monitor))
                    else:
                        print('No improvement over previous best {} value, not saving model.'.format(monitor))
                else:
                    save = True
                if save:
                    self.save(model_save_dir=save_dir, saver=saver, tags=save_tags, name=save_name,
                              include_global_step=True, include_last_training_loss=True,
                              include_metrics=(len(self.metric_names) > 0))

            if self.training_loss < self.best_training_loss:
                self.best_training_loss = self.training_loss
            if epoch % eval_frequency == 0 and len(self.metric_values) == len(self.metric_names):
                for i, metric_name in enumerate(self.metric_names):
                    if (metric_name == 'loss') and (self.metric_values[i] < self.best_metric_values[i]):
                        self.best_metric_values[i] = self.metric_values[i]
                    elif (metric_name in ['accuracy', 'mean_iou']) and (self.metric_values[i] > self.best_metric_values[i]):
                        self.best_metric_values[i] = self.metric_values[i]
        for w_ in (training_writer, evaluation_writer):
            if w_ is not None:
                w_.close()

    def train_on_batch(self, batch_images, batch_labels, learning_rate, keep_prob=0.5, l2_regularization=0.0):
        """One training `sess.run` (fcn8s_tensorflow.py:565-572): H2D feed, forward, backward, Adam, loss fetch."""
        x = self._images_to_device(batch_images)
        y = self._labels_to_device(batch_labels)
        self.engine.train_step(x, y, learning_rate, keep_prob, l2_regularization)
        self.g_step = self.engine.global_step
        return self.engine.loss_value(x.shape)

    # ------------------------------------------------------------------ evaluate (fcn8s_tensorflow.py:660-741)
    def _evaluate(self, data_generator, metrics, num_batches, l2_regularization, description='Running evaluation'):
        from tqdm import trange
        self._conf.zero_()                       # metrics_reset_op, :674
        loss_sum, loss_count = 0.0, 0
        tr = trange(num_batches, file=sys.stdout)
        tr.set_description(description)
        from .feed import Feeder
        feeder = Feeder(self, data_generator, num_batches)
        pending = None
        try:
            for step in tr:
                x, y = feeder.get()
                self.engine.eval_step(x, y, self._conf, l2_regularization)
                feeder.release()
                if 'loss' in self.metric_names:   # tf.metrics.mean over per-batch total_loss, :284 (read one step late)
                    token = self._loss_read_begin(x.shape)
                    if pending is not None:
                        loss_sum += self._loss_read_end(pending)
                    pending = token
                    loss_count += 1
            if pending is not None:
                loss_sum += self._loss_read_end(pending)
        finally:
            feeder.close()
        if self.engine.world > 1:
            # data parallel: every rank evaluated its own shard; the metrics (and every save-best decision taken from
            # them) are those of the whole evaluation set on every rank
            from . import dist as _dist
            pair = torch.tensor([loss_sum, float(loss_count)], dtype=torch.float64, device=self.engine.device)
            _dist.all_reduce_metrics(pair, self._conf)
            loss_sum, loss_count = float(pair[0]), int(round(float(pair[1])))
        cm = self._conf.cpu().numpy()
        self.metric_values = self._metric_values_from(loss_sum, loss_count, cm)
        evaluation_results_string = ''
        for i, metric_name in enumerate(self.metric_names):
            evaluation_results_string += metric_name + ': {:.4f}  '.format(self.metric_values[i])
        print(evaluation_results_string)

    def evaluate(self, data_generator, num_batches, metrics={'loss', 'mean_iou', 'accuracy'}, l2_regularization=0.0, dataset='val'):
        for metric in metrics:
            if not metric in ['loss', 'mean_iou', 'accuracy']:
                raise ValueError("{} is not a valid metric. Valid metrics are ['loss', mean_iou', 'accuracy']".format(metric))
        if not dataset in {'train', 'val'}:
            raise ValueError("`dataset` must be either 'train' or 'val'.")
        self._initialize_metrics(metrics)
        self._evaluate(data_generator, metrics, num_batches, l2_regularization, description='Running evaluation')
        self.eval_dataset = 'val' if dataset == 'val' else 'train'

    # ------------------------------------------------------------------ predict (fcn8s_tensorflow.py:743-855)
    def predict(self, images, argmax=True):
        x = self._images_to_device(images)
        if not argmax:
            return self.engine.predict(x, argmax=False).cpu().numpy()
        # the class map leaves the kernel as one byte per pixel, crosses PCIe into a pinned buffer and is widened to
        # the int64 that tf.argmax returns (fcn8s_tensorflow.py:269,768) on the host: 1/8 of the D2H bytes
        seg = self.engine.predict(x, argmax=True, compact=True)
        k = ("argmax", tuple(seg.shape))
        pin = self._stage.get(k)
        if pin is None:
            pin = self._stage[k] = torch.empty(seg.shape, dtype=torch.uint8).pin_memory()
        pin.copy_(seg, non_blocking=True)
        torch.cuda.current_stream(self.engine.device).synchronize()
        return pin.numpy().astype(np.int64)

    def predict_and_save(self, results_dir, images_dir, color_map, resize=False, image_file_extension='png',
                         include_unprocessed_image=False, arrangement='vertical', overwrite_existing=True):
        from tqdm import trange
        from PIL import Image
        if overwrite_existing and os.path.exists(results_dir):
            shutil.rmtree(results_dir)
        os.makedirs(results_dir)
        image_paths = glob(os.path.join(images_dir, '*.' + image_file_extension))
        print('The segmented images will be saved to "{}"'.format(results_dir))
        tr = trange(len(image_paths), file=sys.stdout)
        tr.set_description('Processing images')
        for i in tr:
            filepath = image_paths[i]
            img = Image.open(filepath).convert('RGB')
            if resize and not np.array_equal((img.height, img.width), resize):
                img = img.resize((resize[1], resize[0]), Image.BILINEAR)
            image = np.asarray(img, dtype=np.uint8)
            h, w, _ = image.shape
            seg = self.predict([image], argmax=True)[0]
            processed = print_segmentation_onto_image(image, seg, color_map)
            if include_unprocessed_image:
                axis = 0 if arrangement == 'vertical' else 1
                processed = np.concatenate([processed, image], axis=axis)
            Image.fromarray(processed).save(os.path.join(results_dir, os.path.basename(filepath)))

    # ------------------------------------------------------------------ persistence (fcn8s_tensorflow.py:857-944)
    def save(self, model_save_dir, saver, tags=['default'], name=None, include_global_step=True,
             include_last_training_loss=True, include_metrics=True, force_save=False):
        if (not self.variables_updated) and (not force_save):
            print("Abort: Nothing to save, no training has been performed since the model was last saved.")
            return
        if not saver in {'saved_model', 'train_saver'}:
            raise ValueError("Unexpected value for `saver`: Can be either 'saved_model' or 'train_saver', but received '{}'.".format(saver))
        if self.training_loss is None:
            include_last_training_loss = False
        model_name = 'saved_model'
        if not name is None:
            model_name += '_' + name
        if include_global_step:
            self.g_step = self.engine.global_step
            model_name += '_(globalstep-{})'.format(self.g_step)
        if include_last_training_loss:
            model_name += '_(trainloss-{:.4f})'.format(self.training_loss)
        if include_metrics:
            if self.eval_dataset == 'val':
                model_name += '_(eval_on_val_dataset)'
            else:
                model_name += '_(eval_on_train_dataset)'
            for i in range(min(len(self.metric_names), len(self.metric_values))):
                model_name += '_({}-{:.4f})'.format(self.metric_names[i], self.metric_values[i])
        if not (include_global_step or include_last_training_loss or include_metrics) and (name is None):
            model_name += '_{}'.format(time.time())
        out_dir = os.path.join(model_save_dir, model_name)
        os.makedirs(out_dir, exist_ok=True)
        e = self.engine
        arrays = {n: t.numpy() for n, t in e.state_dict().items()}
        for n in e.layout:   # Adam slots are global variables in the reference and are saved with the model
            arrays[n + "/Adam"] = e.view(n, e.adam_m).cpu().numpy()
            arrays[n + "/Adam_1"] = e.view(n, e.adam_v).cpu().numpy()
        arrays["optimizer/global_step"] = np.asarray(e.global_step, np.int32)       # tf.Variable(0) at :246
        # tf.train.AdamOptimizer keeps beta^(t+1) in these accumulators after t steps (initialised to beta, multiplied
        # after every step), so a restore in TensorFlow continues with the right bias correction
        arrays["optimizer/beta1_power"] = np.asarray(0.9 ** (e.global_step + 1), np.float32)
        arrays["optimizer/beta2_power"] = np.asarray(0.999 ** (e.global_step + 1), np.float32)
        # The reference writes TensorFlow files: SavedModelBuilder puts the variables under variables/variables.*
        # (:922-925), tf.train.Saver under the prefix <dir>/variables plus a `checkpoint` state file (:926-934).  Both
        # are tensor bundles; tf_bundle.py writes them without TensorFlow (tf.train.load_checkpoint reads them).
        from . import tf_bundle
        if saver == 'saved_model':
            tf_bundle.write_bundle(os.path.join(out_dir, 'variables', 'variables'), arrays)
        else:
            tf_bundle.write_bundle(os.path.join(out_dir, 'variables'), arrays)
            with open(os.path.join(out_dir, 'checkpoint'), 'w') as f:
                f.write('model_checkpoint_path: "variables"\nall_model_checkpoint_paths: "variables"\n')
        self.variables_updated = False
        return out_dir

    def _restore_optimizer(self, data):
        e = self.engine
        for key in ("optimizer/global_step", "global_step"):
            if key in data:
                e.global_step = int(data[key])
                break
        if "optimizer/beta1_power" in data:
            want = 0.9 ** (e.global_step + 1)
            got = float(np.asarray(data["optimizer/beta1_power"]))
            if abs(got - want) > 1e-3 * want:
                warnings.warn("optimizer/beta1_power = %g does not match global_step %d (expected %g); the engine "
                              "derives the Adam bias correction from global_step" % (got, e.global_step, want))
        for n in e.layout:
            if n + "/Adam" in data:
                e.view(n, e.adam_m).copy_(torch.from_numpy(np.asarray(data[n + "/Adam"])).to(e.device))
                e.view(n, e.adam_v).copy_(torch.from_numpy(np.asarray(data[n + "/Adam_1"])).to(e.device))

    def load_variables(self, path):
        data = _load_npz_weights(path, self.num_classes)
        self.engine.load_weights(data)
        self._restore_optimizer(data)

    def close(self):
        self.engine = None
        self._stage = {}
        torch.cuda.empty_cache()
        print("The session has been closed.")
