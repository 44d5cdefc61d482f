"""The drop-in boundary on the GPU: `FCN8s.train / evaluate / predict / save / load` driven by a generator that follows
the reference's protocol (`next(gen)` -> uint8 images [n,H,W,3], bool one-hot labels [n,H,W,C], short last batch;
data_generator/batch_generator.py:244,414-415), checked against the CPU oracle of the reference graph
(fcn8s_tensorflow.py:399-770, 857-944).  Uses the prefetching feed (feed.py) and the CUDA-graph step, i.e. exactly
what a user of the class runs."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

from oracle import fcn8s_oracle as oracle

pytestmark = pytest.mark.gpu
C, H, W = 5, 64, 96


def batches(sizes, seed):
    """Infinite generator cycling over batches of the given sizes (the reference's generators loop for ever)."""
    data = [oracle.synthetic_batch(n, H, W, C, seed=seed + i) for i, n in enumerate(sizes)]
    while True:
        for images, labels in data:
            yield images, labels


def quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


@pytest.fixture(scope="module")
def weights():
    return {k: v.numpy() for k, v in oracle.init_weights(C, seed=4, decoder_std_scale=10.0).items()}


def make_model(cuda_device, weights, precision="fp32"):
    from fcn8s_tensorflow_b200.fcn8s import FCN8s
    return FCN8s(weights=weights, precision=precision, device=cuda_device)


def test_constructor_and_argument_errors_are_the_references(cuda_device, weights):
    from fcn8s_tensorflow_b200.fcn8s import FCN8s
    with pytest.raises(ValueError, match="You must provide either both `model_load_dir` and `tags`"):
        FCN8s()                                                      # fcn8s_tensorflow.py:40-41
    m = make_model(cuda_device, weights)
    gen = batches([2], 0)
    with pytest.raises(ValueError, match="`eval_dataset` must be one of 'train' or 'val'"):      # :511-512
        m.train(gen, 1, 1, lambda s: 1e-4, eval_dataset='test')
    with pytest.raises(ValueError, match="When eval_dataset == 'val'"):                            # :514-515
        m.train(gen, 1, 1, lambda s: 1e-4, eval_dataset='val')
    with pytest.raises(ValueError, match="is not a valid metric"):                                 # :518-519
        m.train(gen, 1, 1, lambda s: 1e-4, metrics={'iou'})
    with pytest.raises(ValueError, match="You are trying to monitor"):                             # :521-522
        m.train(gen, 1, 1, lambda s: 1e-4, metrics={'loss'}, monitor='accuracy')
    with pytest.raises(ValueError, match="`dataset` must be either 'train' or 'val'"):             # :731-732
        m.evaluate(gen, 1, dataset='test')
    with pytest.raises(ValueError, match="Unexpected value for `saver`"):                          # :898-899
        m.variables_updated = True
        m.save("/tmp/nowhere", saver='pickle')


def test_train_evaluate_predict_save_load_roundtrip(cuda_device, weights, tmp_path):
    m = make_model(cuda_device, weights)
    gen = batches([2, 2, 1], 10)          # third batch is short, like the last batch of a pass
    lr = lambda step: 1e-4 if step < 4 else 5e-5   # noqa: E731  (learning_rate_schedule: global step -> lr, :403,527,583)
    quiet(m.train, gen, epochs=2, steps_per_epoch=3, learning_rate_schedule=lr, keep_prob=1.0, l2_regularization=0.01,
          eval_dataset='train', eval_frequency=1, metrics={'loss', 'mean_iou', 'accuracy'},
          save_during_training=True, save_dir=str(tmp_path), save_best_only=False, save_frequency=2,
          record_summaries=False)
    assert m.g_step == m.engine.global_step == 6 and m.variables_updated is False
    assert m.metric_names == ['loss', 'mean_iou', 'accuracy'] and len(m.metric_values) == 3
    assert np.isfinite(m.training_loss) and m.best_training_loss <= 99999999.0
    # directory name scheme of fcn8s_tensorflow.py:904-920 (save_name='' gives the reference's double underscore)
    saved = [d for d in os.listdir(tmp_path) if d.startswith('saved_model__(globalstep-6)_(trainloss-')]
    assert len(saved) == 1 and '_(eval_on_train_dataset)_(loss-' in saved[0] and '_(mean_iou-' in saved[0]

    # the same six steps in the CPU oracle (TF-form Adam, keep_prob 1, l2 0.01), fp64
    w = {k: torch.from_numpy(v.copy()).double() for k, v in weights.items()}
    am = {k: torch.zeros_like(v) for k, v in w.items()}
    av = {k: torch.zeros_like(v) for k, v in w.items()}
    ref_gen = batches([2, 2, 1], 10)
    step = 0
    for _ in range(6):
        images, labels = next(ref_gen)
        _, step = oracle.train_step(w, am, av, step, images, labels, lr(step), keep_prob=1.0, l2_rate=0.01,
                                    dtype=torch.float64)
    # parameters after six steps: Adam moves every weight by ~lr per step whatever the gradient's size, so compare
    # the UPDATE with the oracle's update.  Weights whose gradient is ~0 (most of fc6 at this tiny size) get a full
    # +-lr step whose SIGN is decided inside fp32 noise; tests/test_gpu_engine.py::test_two_adam_steps_match_tf_form
    # allows 0.1 per tensor after 2 steps for the same reason (measured here: 0.09 after 6 steps)
    got = m.engine.state_dict()
    num = sum(float((got[k].double() - w[k]).pow(2).sum()) for k in w)
    den = sum(float((torch.from_numpy(weights[k]).double() - w[k]).pow(2).sum()) for k in w)
    assert (num / den) ** 0.5 <= 0.2, (num / den) ** 0.5

    # evaluate(): streaming mean of per-batch total_loss, mean IoU and pixel accuracy over 3 batches (:284-301)
    quiet(m.evaluate, batches([2, 2, 1], 10), num_batches=3, metrics={'loss', 'mean_iou', 'accuracy'},
          l2_regularization=0.01, dataset='val')
    wt = {k: v.float() for k, v in w.items()}
    cm = np.zeros((C, C), np.int64)
    losses = []
    eg = batches([2, 2, 1], 10)
    for _ in range(3):
        images, labels = next(eg)
        logits = oracle.forward(w, images, dtype=torch.float64)
        losses.append(float(oracle.loss_from_logits(w, logits, labels, 0.01)))
        cm += oracle.confusion_matrix(labels, logits.argmax(-1).numpy(), C)
    vals = dict(zip(m.metric_names, m.metric_values))
    assert m.eval_dataset == 'val'
    assert abs(vals['loss'] - np.mean(losses)) <= 2e-3 * abs(np.mean(losses))
    assert abs(vals['accuracy'] - oracle.accuracy_from_cm(cm)) <= 5e-3
    assert abs(vals['mean_iou'] - oracle.mean_iou_from_cm(cm)) <= 5e-3
    del wt

    # predict(): list of HxWx3 arrays -> int64 [N,H,W]; softmax rows sum to one (:743-770)
    images, _ = oracle.synthetic_batch(2, H, W, C, seed=99)
    pred = m.predict([images[0], images[1]], argmax=True)
    assert pred.dtype == np.int64 and pred.shape == (2, H, W)
    sm = m.predict(images, argmax=False)
    assert sm.dtype == np.float32 and sm.shape == (2, H, W, C) and np.allclose(sm.sum(-1), 1.0, atol=1e-5)
    assert (sm.argmax(-1) == pred).mean() > 0.999

    # save -> load in a new object: same predictions, same global step, Adam slots restored (:72-101, 857-944)
    from fcn8s_tensorflow_b200.fcn8s import FCN8s
    m2 = FCN8s(model_load_dir=os.path.join(str(tmp_path), saved[0]), tags=['default'], precision="fp32",
               device=cuda_device)
    assert m2.engine.global_step == 6 and m2.num_classes == C
    assert np.array_equal(m2.predict(images, argmax=True), pred)
    assert torch.equal(m2.engine.adam_v.cpu(), m.engine.adam_v.cpu())
    b_images, b_labels = oracle.synthetic_batch(2, H, W, C, seed=5)
    l1 = m.train_on_batch(b_images, b_labels, 1e-4, keep_prob=1.0)
    l2 = m2.train_on_batch(b_images, b_labels, 1e-4, keep_prob=1.0)
    assert abs(l1 - l2) <= 1e-5 * abs(l1) and m.g_step == m2.g_step == 7
    quiet(m.close)
    quiet(m2.close)


def test_generator_is_consumed_exactly_like_the_reference(cuda_device, weights):
    """The prefetching feed must pull exactly steps_per_epoch (+ num_batches for evaluation) batches, in order, from a
    generator that training and evaluation share (eval_dataset='train', fcn8s_tensorflow.py:589-604)."""
    m = make_model(cuda_device, weights, precision="bf16")
    pulled = []

    def counting():
        i = 0
        src = batches([2], 3)
        while True:
            pulled.append(i)
            i += 1
            yield next(src)

    quiet(m.train, counting(), epochs=2, steps_per_epoch=2, learning_rate_schedule=lambda s: 1e-4, keep_prob=0.5,
          eval_dataset='train', eval_frequency=1, metrics={'loss'}, record_summaries=False)
    assert pulled == list(range(8))     # 2 epochs x (2 train + 2 eval) batches, nothing prefetched beyond that
    quiet(m.close)


def test_training_and_evaluation_summaries_are_the_references(cuda_device, weights, tmp_path):
    """`record_summaries=True` (fcn8s_tensorflow.py:531-563, 606-608): `total_loss`, `learning_rate` and the variable
    summaries of `_build_summary_ops` (:331-350) every `summaries_frequency` steps; `mean_loss` / `mean_iou` /
    `accuracy` after each evaluation."""
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    from fcn8s_tensorflow_b200.summaries import SUMMARY_VARIABLES
    m = make_model(cuda_device, weights, precision="bf16")
    quiet(m.train, batches([2], 7), epochs=1, steps_per_epoch=3, learning_rate_schedule=lambda s: 1e-4, keep_prob=1.0,
          eval_dataset='train', eval_frequency=1, metrics={'loss', 'mean_iou', 'accuracy'}, record_summaries=True,
          summaries_frequency=2, summaries_dir=str(tmp_path), summaries_name='run')
    tr = EventAccumulator(os.path.join(str(tmp_path), 'run'), size_guidance={"histograms": 0, "scalars": 0})
    tr.Reload()
    tags = tr.Tags()
    assert [e.step for e in tr.Scalars('total_loss')] == [1, 3] and len(tr.Scalars('learning_rate')) == 2
    for _, scope in SUMMARY_VARIABLES:
        assert scope + '/mean' in tags['scalars'] and scope + '/histogram' in tags['histograms']
    w = m.engine.state_dict()['fc6/weights'].double()
    ev = tr.Scalars('fc6/kernel/stddev')[-1]
    assert ev.step == 3 and abs(ev.value - float((w - w.mean()).square().mean().sqrt())) <= 1e-4 * abs(ev.value)
    hv = tr.Histograms('fc6/kernel/histogram')[-1].histogram_value
    assert hv.num == w.numel() and abs(hv.min - float(w.min())) <= 1e-6
    ev_ = EventAccumulator(os.path.join(str(tmp_path), 'run_eval'))
    ev_.Reload()
    assert set(ev_.Tags()['scalars']) == {'mean_loss', 'mean_iou', 'accuracy'}
    quiet(m.close)


def test_feed_ships_class_ids_and_restores_the_one_hot_labels_on_the_device(cuda_device):
    """Labels with >= 8 classes cross PCIe as one id per pixel (fcn8_pack_labels on the host, fcn8_expand_labels on the
    device): what arrives must be the generator's one-hot batch bit for bit (helpers/ground_truth_conversion_utils.py:
    84-88 -> fcn8s_tensorflow.py:559); a batch that is not exactly one-hot travels unchanged."""
    from fcn8s_tensorflow_b200 import ops
    from fcn8s_tensorflow_b200.feed import Feeder
    from fcn8s_tensorflow_b200.fcn8s import check_images, check_labels

    class Stub:
        class engine:
            device = torch.device(cuda_device)
        _check_images = staticmethod(check_images)

        def __init__(self, Cn):
            self.Cn = Cn

        def _check_labels(self, labels):
            return check_labels(labels, self.Cn)

    rng = np.random.default_rng(11)
    for Cn, n, h, w in ((20, 2, 32, 64), (9, 3, 17, 5), (8, 1, 8, 8)):
        ids = rng.integers(0, Cn, size=(n, h, w))
        onehot = np.eye(Cn, dtype=bool)[ids]
        soft = onehot.copy()
        soft[0, 0, 0, :] = False           # a void pixel without any class: not one-hot, must travel as it is
        images = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        # device kernel alone
        d_ids = torch.from_numpy(ids.astype(np.uint8)).to(cuda_device)
        d_out = torch.full(onehot.shape, 7, dtype=torch.uint8, device=cuda_device)
        ops.expand_labels(d_ids, d_out)
        assert np.array_equal(d_out.cpu().numpy(), onehot.view(np.uint8))
        # through the feeder
        stub = Stub(Cn)
        feeder = Feeder(stub, iter([(images, onehot), (images, soft), (images, onehot)]), 3)
        try:
            for expect in (onehot, soft, onehot):
                x, y = feeder.get()
                torch.cuda.current_stream().synchronize()
                assert np.array_equal(x.cpu().numpy(), images)
                assert np.array_equal(y.cpu().numpy(), expect.view(np.uint8))
                feeder.release()
            feeder.thread.join(timeout=30)
            # the last batch staged was a packed one: images + one byte per pixel crossed PCIe
            assert stub.feed_h2d_bytes_per_batch == images.nbytes + ids.size
        finally:
            feeder.close()
