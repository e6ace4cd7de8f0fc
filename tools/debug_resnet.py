"""GPU debug: run the ResNet-50 command list eagerly at a given batch/image and report the first command whose output
contains a non-finite value, plus the loss after each of a few steps."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccv_b200 import nnc, resnet50

batch, image, lr = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]) if len(sys.argv) > 3 else 0.04
nnc.init()
stream = nnc.Stream(0)
net = resnet50.Net(batch, image=image, classes=1000, learn_rate=lr)
rs = np.random.RandomState(1)
net.input.upload(rs.rand(batch, image, image, 3).astype(np.float32))
net.labels.upload(rs.randint(0, 1000, size=(batch,)).astype(np.int32))


def check(nodes, tag):
    for i, (cmd, hint, flags, ins, outs) in enumerate(nodes):
        st = nnc.cmd_exec(cmd, hint, flags, ins, outs, stream)
        stream.wait()
        assert st == 0, (tag, i, hex(cmd.cmd), st)
        for j, t in enumerate(outs):
            if t is None or t.params.datatype != nnc.CCV_32F:
                continue
            a = t.download()
            if not np.isfinite(a).all():
                print("%s node %d cmd 0x%08x output %d dims %s: %d non-finite of %d; max|finite| %.3e" % (tag, i, cmd.cmd, j, t.dims, int((~np.isfinite(a)).sum()), a.size, float(np.abs(a[np.isfinite(a)]).max()) if np.isfinite(a).any() else -1))
                for k, tin in enumerate(ins):
                    if tin is not None and tin.params.datatype == nnc.CCV_32F:
                        b = tin.download()
                        print("   input %d dims %s finite=%s absmax=%.3e" % (k, tin.dims, bool(np.isfinite(b).all()), float(np.abs(b[np.isfinite(b)]).max())))
                return False
    return True


ok = check(net.fwd, "fwd")
print("forward finite:", ok, "loss mean", float(net.loss.download().mean()), "min prob at label", float(np.exp(-net.loss.download()).min()))
if ok:
    ok = check(net.bwd, "bwd")
    print("backward finite:", ok)
if ok:
    g = nnc.Graph()
    for n in net.fwd + net.bwd + net.opt:
        g.exec_new(*n)
    for step in range(6):
        assert g.run(stream) == 0
        stream.wait()
        l = net.loss.download()
        print("step", step, "loss mean %.4f" % float(l.mean()), "finite", bool(np.isfinite(l).all()))
