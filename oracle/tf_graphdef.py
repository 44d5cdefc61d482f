"""The reference's inference graph as a TensorFlow GraphDef -- TEST INFRASTRUCTURE ONLY (like the rest of oracle/).

TensorFlow itself cannot run here (oracle/fcn8s_oracle.py header), but its graph FORMAT can be written without it: the
protobuf classes generated from TensorFlow's own .proto files ship in the `tensorboard` wheel.  This module writes, node
by node, the graph that `fcn8s_tensorflow.py` builds for prediction -- the convolutionalised VGG-16 encoder [EXT] and
the decoder of `_build_decoder` (:154-237) with the ops `tf.multiply`, `tf.layers.conv2d` (Conv2D + BiasAdd),
`tf.layers.conv2d_transpose` (Conv2DBackpropInput + BiasAdd; filter layout [kh, kw, out, in]), `tf.add` and
`tf.nn.softmax` (:268) emit, all NHWC / padding SAME -- so that an INDEPENDENT implementation of TensorFlow's op
semantics can execute it: OpenCV's TensorFlow importer (`cv2.dnn.readNetFromTensorflow`, a C++ code base that shares
nothing with this repository or with PyTorch's convolution code).  `tests/golden/make_golden.py` runs it and commits
the outputs (tests/golden/opencv_tf_fcn8s.npz); `tests/test_oracle.py` checks the oracle against them.

What this pins: SAME padding of the 3x3 / 7x7 / 1x1 convolutions, SAME max-pooling (odd sizes included), the output
alignment and kernel layout of the SAME transposed convolutions (4x4 / 2 and 16x16 / 8), the 1e-4 / 1e-2 skip scales,
the two adds and the softmax -- i.e. the forward arithmetic.  What it does not: TensorFlow's own kernels (still not
run), the loss / optimiser (OpenCV is inference-only; those stay pinned by oracle/tf_semantics.py).
The encoder is built with narrow layers (same topology, fewer channels) to keep the fixture small; the oracle's
forward() takes its widths from the weight shapes.
"""
from collections import OrderedDict

import numpy as np

VGG_TOPOLOGY = [(1, 2), (2, 2), (3, 3), (4, 3), (5, 3)]      # (block, convs): conv1_1 ... conv5_3


def _protos():
    from tensorboard.compat.proto import attr_value_pb2, graph_pb2, tensor_pb2, tensor_shape_pb2, types_pb2
    return attr_value_pb2, graph_pb2, tensor_pb2, tensor_shape_pb2, types_pb2


class Graph:
    """Minimal GraphDef writer: the handful of ops of the reference's graph, with TensorFlow's attribute names."""

    def __init__(self):
        self.av, gp, self.tp, self.sp, self.ty = _protos()
        self.g = gp.GraphDef()
        self.A = self.av.AttrValue
        self.F32, self.I32 = self.ty.DT_FLOAT, self.ty.DT_INT32

    def node(self, name, op, inputs=(), **attrs):
        n = self.g.node.add()
        n.name, n.op = name, op
        n.input.extend(inputs)
        for k, v in attrs.items():
            n.attr[k].CopyFrom(v)
        return name

    def _ints(self, values):
        return self.A(list=self.A.ListValue(i=list(values)))

    def const(self, name, a):
        a = np.ascontiguousarray(a)
        t = self.tp.TensorProto()
        t.dtype = self.F32 if a.dtype == np.float32 else self.I32
        for d in a.shape:
            t.tensor_shape.dim.add().size = d
        t.tensor_content = a.tobytes()
        return self.node(name, "Const", dtype=self.A(type=t.dtype), value=self.A(tensor=t))

    def placeholder(self, name, shape):
        s = self.sp.TensorShapeProto()
        for d in shape:
            s.dim.add().size = d
        return self.node(name, "Placeholder", dtype=self.A(type=self.F32), shape=self.A(shape=s))

    def conv2d(self, name, x, w_hwio, bias, relu, wname="filter", bname="biases"):
        """Conv2D stride 1 SAME + BiasAdd (+ Relu): the encoder layers [EXT] and tf.layers.conv2d (:173-200)."""
        self.const("%s/%s" % (name, wname), w_hwio)
        self.node(name + "/Conv2D", "Conv2D", [x, "%s/%s" % (name, wname)], T=self.A(type=self.F32),
                  strides=self._ints([1, 1, 1, 1]), padding=self.A(s=b"SAME"), data_format=self.A(s=b"NHWC"),
                  dilations=self._ints([1, 1, 1, 1]))
        self.const("%s/%s" % (name, bname), bias)
        out = self.node(name + "/BiasAdd", "BiasAdd", [name + "/Conv2D", "%s/%s" % (name, bname)],
                        T=self.A(type=self.F32), data_format=self.A(s=b"NHWC"))
        if relu:
            out = self.node(name + "/Relu", "Relu", [out], T=self.A(type=self.F32))
        return out

    def max_pool(self, name, x):
        return self.node(name, "MaxPool", [x], T=self.A(type=self.F32), ksize=self._ints([1, 2, 2, 1]),
                         strides=self._ints([1, 2, 2, 1]), padding=self.A(s=b"SAME"), data_format=self.A(s=b"NHWC"))

    def conv2d_transpose(self, name, x, t_hwoi, bias, stride, out_shape):
        """tf.layers.conv2d_transpose(kernel 2s, strides s, padding 'same') (:204-233) = Conv2DBackpropInput(output
        shape, kernel [kh, kw, out, in], x) + BiasAdd."""
        self.const(name + "/output_shape", np.asarray(out_shape, np.int32))
        self.const(name + "/kernel", t_hwoi)
        self.node(name + "/conv2d_transpose", "Conv2DBackpropInput", [name + "/output_shape", name + "/kernel", x],
                  T=self.A(type=self.F32), strides=self._ints([1, stride, stride, 1]), padding=self.A(s=b"SAME"),
                  data_format=self.A(s=b"NHWC"), dilations=self._ints([1, 1, 1, 1]))
        self.const(name + "/bias", bias)
        return self.node(name + "/BiasAdd", "BiasAdd", [name + "/conv2d_transpose", name + "/bias"],
                         T=self.A(type=self.F32), data_format=self.A(s=b"NHWC"))

    def binary(self, name, op, a, b):
        return self.node(name, op, [a, b], T=self.A(type=self.F32))

    def softmax(self, name, x):
        return self.node(name, "Softmax", [x], T=self.A(type=self.F32))

    def serialized(self):
        """The GraphDef as a uint8 array (what cv2.dnn.readNetFromTensorflow takes as an in-memory model)."""
        return np.frombuffer(self.g.SerializeToString(), np.uint8)


def narrow_weights(num_classes, widths=(4, 8, 16, 32, 32), fc=64, seed=0):
    """Seeded weights of an FCN-8s with the reference's topology and variable names but narrow layers, O(1) logits."""
    rng = np.random.default_rng(seed)
    w = OrderedDict()
    cin = 3
    for (b, n), cout in zip(VGG_TOPOLOGY, widths):
        for i in range(1, n + 1):
            w["conv%d_%d/filter" % (b, i)] = (rng.standard_normal((3, 3, cin, cout)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
            w["conv%d_%d/biases" % (b, i)] = (rng.standard_normal(cout) * 0.1).astype(np.float32)
            cin = cout
    w["fc6/weights"] = (rng.standard_normal((7, 7, cin, fc)) * np.sqrt(2.0 / (49 * cin))).astype(np.float32)
    w["fc6/biases"] = (rng.standard_normal(fc) * 0.1).astype(np.float32)
    w["fc7/weights"] = (rng.standard_normal((1, 1, fc, fc)) * np.sqrt(2.0 / fc)).astype(np.float32)
    w["fc7/biases"] = (rng.standard_normal(fc) * 0.1).astype(np.float32)
    for name, ci, std in (("pool3_1x1", widths[2], 3.0), ("pool4_1x1", widths[3], 3.0), ("fc7_1x1", fc, 0.1)):
        w[name + "/kernel"] = (rng.standard_normal((1, 1, ci, num_classes)) * std).astype(np.float32)
        w[name + "/bias"] = (rng.standard_normal(num_classes) * 0.1).astype(np.float32)
    for name, k in (("fc7_conv2d_trans", 4), ("fc7_pool4_conv2d_trans", 4), ("fc7_pool4_pool3_conv2d_trans", 16)):
        w[name + "/kernel"] = (rng.standard_normal((k, k, num_classes, num_classes)) * 0.3).astype(np.float32)
        w[name + "/bias"] = (rng.standard_normal(num_classes) * 0.1).astype(np.float32)
    return w


def inference_graph(weights, n, height, width, num_classes):
    """GraphDef of the reference's prediction path on a pre-processed (BGR - mean) NHWC float input.
    Returns (Graph, names of the tensors worth comparing)."""
    g = Graph()
    x = g.placeholder("image_input_preprocessed", [n, height, width, 3])
    pools = {}
    for b, convs in VGG_TOPOLOGY:
        for i in range(1, convs + 1):
            name = "conv%d_%d" % (b, i)
            x = g.conv2d(name, x, weights[name + "/filter"], weights[name + "/biases"], True)
        x = pools[b] = g.max_pool("pool%d" % b, x)
    x = g.conv2d("fc6", x, weights["fc6/weights"], weights["fc6/biases"], True, wname="weights")
    x = g.conv2d("fc7", x, weights["fc7/weights"], weights["fc7/biases"], True, wname="weights")
    # decoder: fcn8s_tensorflow.py:164-235
    g.const("pool3_scale", np.float32(0.0001))
    g.const("pool4_scale", np.float32(0.01))
    p3 = g.binary("pool3_out_scaled", "Mul", pools[3], "pool3_scale")                                     # :171
    s3 = g.conv2d("pool3_1x1", p3, weights["pool3_1x1/kernel"], weights["pool3_1x1/bias"], False, "kernel", "bias")
    p4 = g.binary("pool4_out_scaled", "Mul", pools[4], "pool4_scale")                                     # :182
    s4 = g.conv2d("pool4_1x1", p4, weights["pool4_1x1/kernel"], weights["pool4_1x1/bias"], False, "kernel", "bias")
    s7 = g.conv2d("fc7_1x1", x, weights["fc7_1x1/kernel"], weights["fc7_1x1/bias"], False, "kernel", "bias")
    u2 = g.conv2d_transpose("fc7_conv2d_trans", s7, weights["fc7_conv2d_trans/kernel"], weights["fc7_conv2d_trans/bias"],
                            2, [n, height // 16, width // 16, num_classes])                               # :204-211
    f4 = g.binary("add_fc7_pool4", "Add", u2, s4)                                                          # :213
    u4 = g.conv2d_transpose("fc7_pool4_conv2d_trans", f4, weights["fc7_pool4_conv2d_trans/kernel"],
                            weights["fc7_pool4_conv2d_trans/bias"], 2, [n, height // 8, width // 8, num_classes])
    f3 = g.binary("add_fc7_pool4_pool3", "Add", u4, s3)                                                    # :224
    g.conv2d_transpose("fc7_pool4_pool3_conv2d_trans", f3, weights["fc7_pool4_pool3_conv2d_trans/kernel"],
                       weights["fc7_pool4_pool3_conv2d_trans/bias"], 8, [n, height, width, num_classes])  # :226-235
    g.softmax("softmax_output", "fc7_pool4_pool3_conv2d_trans/BiasAdd")                                    # :268
    # OpenCV fuses BiasAdd into the preceding layer and names the result after it
    outputs = OrderedDict(pool3="pool3", pool4="pool4", fc7="fc7/Relu", f4="add_fc7_pool4", f3="add_fc7_pool4_pool3",
                          logits="fc7_pool4_pool3_conv2d_trans/conv2d_transpose", softmax="softmax_output")
    return g, outputs


def op_graph(x_shape, w_hwio, bias, t_hwoi, t_bias, stride):
    """conv 3x3 SAME + bias -> max-pool SAME -> transposed conv SAME on an ODD-sized input (the sizes the aligned
    full graph never produces).  Returns (Graph, outputs)."""
    n, h, w, _ = x_shape
    g = Graph()
    x = g.placeholder("x", list(x_shape))
    c = g.conv2d("conv", x, w_hwio, bias, False)
    p = g.max_pool("pool", c)
    ho, wo = (h + 1) // 2, (w + 1) // 2
    g.conv2d_transpose("up", p, t_hwoi, t_bias, stride, [n, ho * stride, wo * stride, t_hwoi.shape[2]])
    return g, OrderedDict(conv="conv/Conv2D", pool="pool", up="up/conv2d_transpose")


def run_opencv(graph, outputs, x_nhwc):
    """Execute a Graph with OpenCV's TensorFlow importer; returns name -> NHWC float32 array."""
    import cv2
    net = cv2.dnn.readNetFromTensorflow(graph.serialized())
    net.setInput(np.ascontiguousarray(np.asarray(x_nhwc, np.float32).transpose(0, 3, 1, 2)))    # OpenCV blobs are NCHW
    res = net.forward(list(outputs.values()))
    return OrderedDict((k, np.ascontiguousarray(r.transpose(0, 2, 3, 1))) for k, r in zip(outputs, res))
