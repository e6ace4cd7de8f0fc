"""ResNet-50 (the v1d variant of the reference's trainer, bin/nnc/imagenet.c:17-95) as a flat, topologically ordered
list of nnc commands on CCV_NNC_BACKEND_GPU_SM100 -- the same command sequence ccv_cnnp_model_fit ends up issuing
through ccv_nnc_graph_run (lib/nnc/ccv_cnnp_model.c:1533-1570): forward, softmax + categorical cross-entropy,
backward, and one nesterov-SGD command per parameter (bin/nnc/imagenet.c:314-317).

ccv's symbolic graph, autodiff and memory planner are NOT rebuilt here (out of scope); this module
is only the driver that exercises the hot path at BASELINE.json's headline configurations.  NHWC; fp32 (configs[2]) or, with
dtype = CCV_16BF / CCV_16F, the 16-bit form of configs[3] (bin/nnc/imagenet.c:344 trains in CCV_16F): activations, filters,
dense weights, their biases and all their gradients are 16-bit; batch-norm scale / bias / statistics and their gradients stay
fp32 (lib/nnc/ccv_cnnp_model_addons.c:954-956); every parameter has an fp32 master copy and fp32 momentum that the SGD
commands update from the 16-bit gradients (the mixed form of sgd/gpu/ccv_nnc_sgd_gpu_ref.cu:71-74), and ONE
DATATYPE_CONVERSION command per step refreshes the 16-bit working copy of the filters from the masters; softmax and the
cross-entropy loss run in fp32 on the converted logits.

Data parallelism: one process per GPU holds one batch shard and a full replica; all gradients live in ONE
flat buffer so that the exchange step is a single sum-allreduce, after which every replica runs the same SGD commands
(sgd.scale = 1 / global_batch carries the averaging, bin/nnc/imagenet.c:474-475)."""
import numpy as np

from . import abi, nnc


class GpuFactory(object):
    """Allocates ccv_nnc_tensor_t on the GPU through the backend library. (The oracle-side twin, used by the tests and by
    bench.py's cpu_baseline leg to run the identical command list on CCV_NNC_BACKEND_CPU_REF, lives in oracle/.)"""

    def __init__(self, device=0):
        self.device = device

    def alloc(self, shape, datatype=abi.CCV_32F):
        return nnc.gpu_tensor(list(shape), abi.CCV_TENSOR_FORMAT_NHWC, datatype, self.device)

    def alias(self, base, elem_offset, shape, datatype=abi.CCV_32F):
        return nnc.gpu_tensor(list(shape), abi.CCV_TENSOR_FORMAT_NHWC, datatype, self.device, base.data_ptr + elem_offset * abi.DTYPE_SIZE[datatype])


class Net(object):
    def __init__(self, batch, image=224, classes=1000, device=0, global_batch=None, learn_rate=0.04, weight_decay=1e-4, seed=0, algorithm=-1, factory=None, dtype=abi.CCV_32F):
        self.batch, self.image, self.classes, self.device = batch, image, classes, device
        self.dtype = dtype                      # activations / filters: CCV_32F, CCV_16BF or CCV_16F
        self.esz = abi.DTYPE_SIZE[dtype]
        self.factory = factory if factory is not None else GpuFactory(device)
        self.global_batch = global_batch or batch
        self.learn_rate, self.weight_decay = learn_rate, weight_decay
        self.algorithm = algorithm
        self.rng = np.random.RandomState(seed)
        self.tensors = []       # everything allocated (for free())
        self.params = []        # (name, shape, decay) in creation order
        self.param_init = {}
        self.fwd, self.bwd, self.opt = [], [], []   # node lists: (cmd, hint, flags, inputs, outputs)
        self.flops = 0
        self.bytes_alloc = 0
        self._pending_params = []
        self._build()

    # ------------------------------------------------------------------------------------------------ allocation
    def alloc(self, shape, datatype=abi.CCV_32F):
        t = self.factory.alloc(shape, datatype)
        self.tensors.append(t)
        self.bytes_alloc += int(np.prod(shape)) * abi.DTYPE_SIZE[datatype]
        return t

    def act(self, shape):
        """an activation (or activation gradient) tensor: the model's dtype"""
        return self.alloc(shape, self.dtype)

    def alias(self, base, elem_offset, shape, datatype=abi.CCV_32F):
        t = self.factory.alias(base, elem_offset, shape, datatype) if datatype != abi.CCV_32F else self.factory.alias(base, elem_offset, shape)
        self.tensors.append(t)
        return t

    def param(self, name, shape, init, decay, low_precision=True):
        """Parameters, their gradients and momenta are carved out of flat buffers once all shapes are known.  low_precision:
        in a 16-bit model this parameter is used (and its gradient produced) in the 16-bit type; False = fp32 throughout
        (batch-norm scale / bias)."""
        self._pending_params.append((name, tuple(shape), init, decay, low_precision and self.dtype != abi.CCV_32F))
        return len(self._pending_params) - 1

    def _materialise_params(self):
        """fp32 model: three flat fp32 buffers (parameters w_flat, gradients g_flat, momenta m_flat) so that the data-parallel
        exchange is ONE allreduce.  16-bit model: the same three for the fp32-only parameters (batch norm) as w_flat_b / g_flat_b /
        m_flat_b, and for the 16-bit group fp32 masters w_flat + fp32 momenta m_flat + a 16-bit working copy w16_flat + 16-bit
        gradients g_flat (what the allreduce moves: half the bytes)."""
        mixed = self.dtype != abi.CCV_32F
        offs, total = [], {False: 0, True: 0}
        for _, shape, _, _, low in self._pending_params:
            offs.append(total[low])
            total[low] += (int(np.prod(shape)) + 63) // 64 * 64  # 256-byte aligned fp32 slices (128-byte for 16-bit): TMA friendly
        self.w, self.g, self.m, self.w_master = [], [], [], []
        self.grad_slices = []   # per parameter: (flat gradient buffer, element offset, padded element count)
        if not mixed:
            self.flat_count = total[False]
            self.w_flat, self.g_flat, self.m_flat = (self.alloc([total[False]]) for _ in range(3))
            host = np.zeros((total[False],), np.float32)
            for (name, shape, init, decay, _), o in zip(self._pending_params, offs):
                n = int(np.prod(shape))
                host[o:o + n] = init.reshape(-1)
                for flat, lst in ((self.w_flat, self.w), (self.g_flat, self.g), (self.m_flat, self.m)):
                    lst.append(self.alias(flat, o, shape))
                self.params.append((name, shape, decay))
                self.grad_slices.append((self.g_flat, o, (n + 63) // 64 * 64))
            self.w_master = self.w
            self.w_flat.upload(host)
            self.m_flat.upload(np.zeros((total[False],), np.float32))
            self.g_flat.upload(np.zeros((total[False],), np.float32))
            self.param_host, self.param_offsets = host, offs
            self.grad_flats = [self.g_flat]
            return
        na, nb = total[True], max(total[False], 64)
        self.flat_count = na
        self.w_flat, self.m_flat = self.alloc([na]), self.alloc([na])                      # fp32 masters and momenta
        self.w16_flat, self.g_flat = self.alloc([na], self.dtype), self.alloc([na], self.dtype)  # 16-bit working copy, 16-bit gradients
        self.w_flat_b, self.g_flat_b, self.m_flat_b = (self.alloc([nb]) for _ in range(3))
        host_a, host_b = np.zeros((na,), np.float32), np.zeros((nb,), np.float32)
        for (name, shape, init, decay, low), o in zip(self._pending_params, offs):
            n = int(np.prod(shape))
            if low:
                host_a[o:o + n] = init.reshape(-1)
                self.w.append(self.alias(self.w16_flat, o, shape, self.dtype))
                self.g.append(self.alias(self.g_flat, o, shape, self.dtype))
                self.w_master.append(self.alias(self.w_flat, o, shape))
                self.m.append(self.alias(self.m_flat, o, shape))
                self.grad_slices.append((self.g_flat, o, (n + 63) // 64 * 64))
            else:
                host_b[o:o + n] = init.reshape(-1)
                self.w.append(self.alias(self.w_flat_b, o, shape))
                self.g.append(self.alias(self.g_flat_b, o, shape))
                self.w_master.append(self.w[-1])
                self.m.append(self.alias(self.m_flat_b, o, shape))
                self.grad_slices.append((self.g_flat_b, o, (n + 63) // 64 * 64))
            self.params.append((name, shape, decay))
        self.w_flat.upload(host_a), self.w_flat_b.upload(host_b)
        for t in (self.m_flat, self.m_flat_b, self.g_flat_b):
            t.upload(np.zeros((t.count,), np.float32))
        for t in (self.w16_flat, self.g_flat):  # all-zero bits are 0.0 in bf16 and fp16 alike; the first forward converts the masters
            t.upload(np.zeros((t.count,), np.uint16))
        self.param_host, self.param_offsets = host_a, offs
        self.grad_flats = [self.g_flat, self.g_flat_b]

    # ------------------------------------------------------------------------------------------------ layers
    def _conv(self, x, xs, cin, cout, k, stride, pad, bias, name, need_dx=True):
        """CONVOLUTION_FORWARD / BACKWARD pair. Returns (y, y_shape)."""
        N, H, W, _ = xs
        P, Q = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        fan_in = cin * k * k
        w = self.param(name + ".w", (cout, k, k, cin), (self.rng.standard_normal((cout, k, k, cin)) * np.sqrt(2.0 / fan_in)).astype(np.float32), self.weight_decay)
        b = self.param(name + ".b", (cout,), np.zeros((cout,), np.float32), 0.0) if bias else None
        y = self.act((N, P, Q, cout))
        self.flops += 2 * N * P * Q * cout * cin * k * k * (3 if need_dx else 2)
        rec = dict(kind="conv", x=x, y=y, w=w, b=b, cin=cin, cout=cout, k=k, hint=abi.hint((stride, stride), (pad, pad)), need_dx=need_dx, xs=xs, name=name)
        self.layers.append(rec)
        return y, (N, P, Q, cout)

    def _bn(self, x, xs, name, relu):
        C = xs[-1]
        scale = self.param(name + ".scale", (1, 1, 1, C), np.ones((1, 1, 1, C), np.float32), 0.0, low_precision=False)
        bias = self.param(name + ".bias", (1, 1, 1, C), np.zeros((1, 1, 1, C), np.float32), 0.0, low_precision=False)
        mean, var = self.alloc((1, 1, 1, C)), self.alloc((1, 1, 1, C))
        mean.upload(np.zeros((C,), np.float32)), var.upload(np.ones((C,), np.float32))
        y = self.act(xs)
        rec = dict(kind="bn", x=x, y=y, scale=scale, bias=bias, mean=mean, var=var, sm=self.alloc((1, 1, 1, C)), sis=self.alloc((1, 1, 1, C)), relu=relu, xs=xs, name=name)
        self.layers.append(rec)
        return y

    def _block(self, x, xs, filters, expansion, stride, projection, name):
        N, H, W, C = xs
        short, short_s = x, xs
        rec_short = []
        start = len(self.layers)
        if projection:
            if stride > 1:
                ps = (N, H // stride, W // stride, C)
                p = self.act(ps)
                self.layers.append(dict(kind="avgpool", x=x, y=p, k=stride, hint=abi.hint((stride, stride), (0, 0)), xs=xs, ys=ps, name=name + ".avgdown"))
                short, short_s = p, ps
            short, short_s = self._conv(short, short_s, C, filters * expansion, 1, 1, 0, False, name + ".conv0")
        branch_end = len(self.layers)
        o, os_ = self._conv(x, xs, C, filters, 1, 1, 0, True, name + ".conv1")
        o = self._bn(o, os_, name + ".bn1", True)
        o, os_ = self._conv(o, os_, filters, filters, 3, stride, 1, True, name + ".conv2")
        o = self._bn(o, os_, name + ".bn2", True)
        o, os_ = self._conv(o, os_, filters, filters * expansion, 1, 1, 0, True, name + ".conv3")
        o = self._bn(o, os_, name + ".bn3", False)
        y = self.act(os_)
        self.layers.append(dict(kind="block_end", main=o, short=short, y=y, x=x, xs=xs, ys=os_, short_range=(start, branch_end), main_range=(branch_end, len(self.layers)), name=name))
        return y, os_

    # ------------------------------------------------------------------------------------------------ build
    def _build(self):
        N, S = self.batch, self.image
        self.layers = []
        self.input = self.act((N, S, S, 3))
        self.labels = self.alloc((N,), datatype=abi.CCV_32S)
        x, xs = self._build_body(self.input, (N, S, S, 3))
        self._build_head(x, xs)

    def _relu(self, x, xs, name):
        """a stand-alone in-place RELU (the ResNet body only has them fused behind a batch norm)"""
        self.layers.append(dict(kind="relu", x=x, y=x, xs=xs, name=name))
        return x

    def _build_body(self, x, xs):
        N = self.batch
        # stem: 3x3/2 (3->32), 3x3 (32->32), 3x3 (32->64), each + BN + ReLU, then 3x3/2 max pool (imagenet.c:73-84)
        for i, (cin, cout, st) in enumerate(((3, 32, 2), (32, 32, 1), (32, 64, 1))):
            x, xs = self._conv(x, xs, cin, cout, 3, st, 1, False, "stem.conv%d" % i, need_dx=i > 0)
            x = self._bn(x, xs, "stem.bn%d" % i, True)
        ps = (N, (xs[1] + 2 - 3) // 2 + 1, (xs[2] + 2 - 3) // 2 + 1, xs[3])
        p = self.act(ps)
        self.layers.append(dict(kind="maxpool", x=x, y=p, k=3, hint=abi.hint((2, 2), (1, 1)), xs=xs, ys=ps, name="stem.pool"))
        x, xs = p, ps
        for li, (filters, stride, blocks) in enumerate(((64, 1, 3), (128, 2, 4), (256, 2, 6), (512, 2, 3))):
            for bi in range(blocks):
                x, xs = self._block(x, xs, filters, 4, stride if bi == 0 else 1, bi == 0, "layer%d.block%d" % (li + 1, bi))
        # global average pool (imagenet.c:89)
        gs = (N, 1, 1, xs[3])
        gp = self.act(gs)
        self.layers.append(dict(kind="avgpool", x=x, y=gp, k=xs[1], hint=abi.hint((1, 1), (0, 0)), xs=xs, ys=gs, name="global_pool"))
        return gp, gs

    def _build_head(self, gp, gs):
        """flatten, dense (+bias), softmax, categorical cross-entropy (imagenet.c:90-93,357)"""
        N = self.batch
        feat = int(np.prod(gs[1:]))
        # ccv_cnnp_dense: w = [out, in] used with transpose_b (lib/nnc/ccv_cnnp_model_addons.c:1360-1373)
        dw = self.param("fc.w", (self.classes, feat), (self.rng.standard_normal((self.classes, feat)) * np.sqrt(1.0 / feat)).astype(np.float32), self.weight_decay)
        db = self.param("fc.b", (self.classes,), np.zeros((self.classes,), np.float32), 0.0)
        self.flops += 2 * N * feat * self.classes * 3
        self._materialise_params()
        self.feat_view = self.alias(gp, 0, (N, feat), self.dtype)  # flatten = alias (ccv_cnnp_flatten)
        self.head_input = gp
        self.logits, self.probs, self.loss = self.alloc((N, self.classes)), self.alloc((N, self.classes)), self.alloc((N,))
        # 16-bit model: the dense layer writes 16-bit logits, softmax + loss run in fp32 on their conversion (bin/nnc/imagenet.c:258-262)
        self.logits_lp = self.act((N, self.classes)) if self.dtype != abi.CCV_32F else self.logits
        self._emit(dw, db)

    def _node(self, lst, cmd, hint, flags, inputs, outputs):
        lst.append((cmd, hint, flags, inputs, outputs))

    def _emit(self, fc_w, fc_b):
        W, G = self.w, self.g
        A = self.algorithm
        f = self.fwd
        mixed = self.dtype != abi.CCV_32F
        if mixed:  # refresh the 16-bit working copy of the filters / dense weights from the fp32 masters: one command per step
            self._node(f, nnc.CMD_DATATYPE_CONVERSION_FORWARD(), None, 0, [self.w_flat], [self.w16_flat])
        for L in self.layers:
            k = L["kind"]
            if k == "conv":
                ins = [L["x"], W[L["w"]]] + ([W[L["b"]]] if L["b"] is not None else [])
                self._node(f, nnc.CMD_CONVOLUTION_FORWARD(1, L["cout"], L["k"], L["k"], L["cin"], algorithm=A), L["hint"], 0, ins, [L["y"]])
            elif k == "bn":
                self._node(f, nnc.CMD_BATCH_NORM_FORWARD(1e-4, 0, 0.9), None, 0, [L["x"], W[L["scale"]], W[L["bias"]], L["mean"], L["var"]], [L["y"], L["mean"], L["var"], L["sm"], L["sis"]])
                if L["relu"]:
                    self._node(f, nnc.CMD_RELU_FORWARD(), None, 0, [L["y"]], [L["y"]])
            elif k == "relu":
                self._node(f, nnc.CMD_RELU_FORWARD(), None, 0, [L["x"]], [L["y"]])
            elif k == "avgpool":
                self._node(f, nnc.CMD_AVERAGE_POOL_FORWARD(L["k"], L["k"]), L["hint"], 0, [L["x"]], [L["y"]])
            elif k == "maxpool":
                self._node(f, nnc.CMD_MAX_POOL_FORWARD(L["k"], L["k"]), L["hint"], 0, [L["x"]], [L["y"]])
            elif k == "block_end":
                self._node(f, nnc.CMD_EWSUM_FORWARD(), None, 0, [L["main"], L["short"]], [L["y"]])
                self._node(f, nnc.CMD_RELU_FORWARD(), None, 0, [L["y"]], [L["y"]])
        self._node(f, nnc.CMD_GEMM_FORWARD((0, 0), (0, 1), algorithm=A), None, 0, [self.feat_view, W[fc_w], W[fc_b]], [self.logits_lp])
        if mixed:
            self._node(f, nnc.CMD_DATATYPE_CONVERSION_FORWARD(), None, 0, [self.logits_lp], [self.logits])
        self._node(f, nnc.CMD_SOFTMAX_FORWARD(), None, 0, [self.logits], [self.probs])
        self._node(f, nnc.CMD_CATEGORICAL_CROSSENTROPY_FORWARD(), None, 0, [self.probs, self.labels], [self.loss])
        # ---------------------------------------------------------------- backward (reverse topological order)
        b = self.bwd
        g_probs, g_logits = self.alloc((self.batch, self.classes)), self.alloc((self.batch, self.classes))
        self._node(b, nnc.CMD_CATEGORICAL_CROSSENTROPY_BACKWARD(), None, 0, [None, self.probs, self.labels], [g_probs])
        self._node(b, nnc.CMD_SOFTMAX_BACKWARD(), None, 0, [g_probs, None, self.probs], [g_logits])
        last = self.layers[-1]
        last_shape = last["ys"] if "ys" in last else last["xs"] if last["kind"] == "relu" else tuple(last["y"].dims)
        g_feat = self.act(last_shape)
        g_feat_view = self.alias(g_feat, 0, (self.batch, int(np.prod(last_shape[1:]))), self.dtype)
        g_logits_lp = g_logits
        if mixed:
            g_logits_lp = self.act((self.batch, self.classes))
            self._node(b, nnc.CMD_DATATYPE_CONVERSION_FORWARD(), None, 0, [g_logits], [g_logits_lp])
        self._node(b, nnc.CMD_GEMM_BACKWARD((0, 0), (0, 1), algorithm=A), None, 0, [g_logits_lp, self.feat_view, W[fc_w]], [g_feat_view, G[fc_w], G[fc_b]])
        grad = {id(self.head_input): g_feat}   # gradient tensor of each activation, keyed by the activation

        def grad_of(t):
            return grad[id(t)]

        def set_grad(t, shape):
            gt = self.act(shape)
            grad[id(t)] = gt
            return gt

        def back_layer(L):
            k = L["kind"]
            if k == "conv":
                gy = grad_of(L["y"])
                gx = set_grad(L["x"], L["xs"]) if L["need_dx"] else None
                outs = [gx, G[L["w"]]] + ([G[L["b"]]] if L["b"] is not None else [])
                self._node(b, nnc.CMD_CONVOLUTION_BACKWARD(1, L["cout"], L["k"], L["k"], L["cin"], algorithm=A), L["hint"], 0, [gy, L["x"], W[L["w"]]], outs)
            elif k == "bn":
                gy = grad_of(L["y"])
                if L["relu"]:
                    self._node(b, nnc.CMD_RELU_BACKWARD(), None, 0, [gy, None, L["y"]], [gy])
                gx = set_grad(L["x"], L["xs"])
                ins = [gy] + [None] * 4 + [L["x"], W[L["scale"]]] + [None] * 6 + [L["sm"], L["sis"]]
                self._node(b, nnc.CMD_BATCH_NORM_BACKWARD(1e-4, 0, 0.9), None, 0, ins, [gx, G[L["scale"]], G[L["bias"]]])
            elif k == "relu":
                gy = grad_of(L["y"])  # in place: the gradient of the input is the masked gradient of the output
                self._node(b, nnc.CMD_RELU_BACKWARD(), None, 0, [gy, None, L["y"]], [gy])
            elif k == "avgpool":
                gx = set_grad(L["x"], L["xs"])
                self._node(b, nnc.CMD_AVERAGE_POOL_BACKWARD(L["k"], L["k"]), L["hint"], 0, [grad_of(L["y"])], [gx])
            elif k == "maxpool":
                gx = set_grad(L["x"], L["xs"])
                self._node(b, nnc.CMD_MAX_POOL_BACKWARD(L["k"], L["k"]), L["hint"], 0, [grad_of(L["y"]), L["x"], L["y"]], [gx])

        i = len(self.layers) - 1
        while i >= 0:
            L = self.layers[i]
            if L["kind"] != "block_end":
                back_layer(L)
                i -= 1
                continue
            gy = grad_of(L["y"])
            self._node(b, nnc.CMD_RELU_BACKWARD(), None, 0, [gy, None, L["y"]], [gy])
            # EWSUM backward hands the same gradient to both branches (ew/ccv_nnc_ew_cpu_ref.c:216-233): alias, no copy
            grad[id(L["main"])] = gy
            grad[id(L["short"])] = gy
            m0, m1 = L["main_range"]
            for j in range(m1 - 1, m0 - 1, -1):
                back_layer(self.layers[j])
            g_main = grad_of(L["x"])
            s0, s1 = L["short_range"]
            if s1 > s0:
                del grad[id(L["x"])]
                for j in range(s1 - 1, s0 - 1, -1):
                    back_layer(self.layers[j])
                g_short = grad_of(L["x"])
            else:
                g_short = gy
            # the block input feeds both branches: its gradient is their sum
            self._node(b, nnc.CMD_EWSUM_FORWARD(), None, 0, [g_main, g_short], [g_main])
            grad[id(L["x"])] = g_main
            i = s0 - 1
        # ---------------------------------------------------------------- optimizer: one SGD command per parameter
        for idx, (name, shape, decay) in enumerate(self.params):
            cmd = nnc.CMD_SGD_FORWARD(1, self.learn_rate, 1.0 / self.global_batch, decay, 0.9, 0.0)
            self._node(self.opt, cmd, None, 0, [G[idx], self.w_master[idx], self.m[idx]], [self.w_master[idx], self.m[idx]])

    # ------------------------------------------------------------------------------------------------ data-parallel exchange
    def backward_with_exchange(self, buckets=4):
        """The backward node list with the gradient exchange woven in (the reference places one allreduce per
        parameter right behind the node that produces its gradient, lib/nnc/ccv_nnc_symbolic_graph_parallel.c:546-575).  The flat
        gradient buffer is cut into `buckets` contiguous slices of similar size, walking the parameters from the last layer
        backwards -- the order in which the backward pass completes them; one COMM_ALLREDUCE_FORWARD node per slice is placed
        behind the first convolution / GEMM backward node after which the whole slice has been written.  Returns
        (nodes, side) where side lists the indices of the allreduce nodes: issued on the graph's side stream
        (Graph.set_side_stream) they run under the remaining backward kernels; the graph's end joins them, so the SGD commands that
        follow see the summed gradients.  A 16-bit model's small fp32 batch-norm gradient buffer travels with the last slice."""
        ready = {}
        for j, (_, _, _, _, outs) in enumerate(self.bwd):
            for t in outs:
                if t is not None:
                    ready[id(t)] = j
        main = [i for i, (flat, _, _) in enumerate(self.grad_slices) if flat is self.g_flat]
        total = sum(self.grad_slices[i][2] for i in main)
        cuts, acc, cur = [], 0, []
        for i in reversed(main):  # parameters are created in forward order: the backward pass finishes them last to first
            cur.append(i)
            acc += self.grad_slices[i][2]
            if acc >= total * (len(cuts) + 1) / float(buckets) and len(cuts) < buckets - 1:
                cuts.append(cur)
                cur = []
        if cur:
            cuts.append(cur)
        inserts = []  # (after node index, [tensors])
        for k, members in enumerate(cuts):
            lo = min(self.grad_slices[i][1] for i in members)
            hi = max(self.grad_slices[i][1] + self.grad_slices[i][2] for i in members)
            done = max(ready[id(self.g[i])] for i in members)
            tensors = [self.alias(self.g_flat, lo, (hi - lo,), self.dtype if self.g_flat.params.datatype != abi.CCV_32F else abi.CCV_32F)]
            if k == len(cuts) - 1:
                if getattr(self, "g_flat_b", None) is not None:
                    tensors.append(self.g_flat_b)
                done = len(self.bwd) - 1
            j = done
            while j < len(self.bwd) - 1 and self.bwd[j][0].cmd not in (abi.CCV_NNC_CONVOLUTION_BACKWARD, abi.CCV_NNC_GEMM_BACKWARD):
                j += 1  # never split a pair the peephole pass fuses (ReLU backward + batch-norm backward, batch norm + convolution)
            inserts.append((j, tensors))
        nodes, side = [], []
        by_pos = {}
        for j, tensors in inserts:
            by_pos.setdefault(j, []).extend(tensors)
        for j, node in enumerate(self.bwd):
            nodes.append(node)
            if j in by_pos:
                side.append(len(nodes))
                nodes.append((nnc.CMD_COMM_ALLREDUCE_FORWARD(), None, 0, by_pos[j], by_pos[j]))
        return nodes, side

    # ------------------------------------------------------------------------------------------------ execution
    def graphs(self):
        """Three flat graphs (forward, backward, optimizer) over the node lists."""
        out = []
        for nodes in (self.fwd, self.bwd, self.opt):
            gph = nnc.Graph()
            for cmd, hint, flags, ins, outs in nodes:
                gph.exec_new(cmd, hint, flags, ins, outs)
            out.append(gph)
        return out

    def free(self):
        for t in self.tensors:
            t.free()
        self.tensors = []
