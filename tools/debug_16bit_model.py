"""GPU debug: run the 16-bit ResNet-50 node by node (unfused, then fused) and report the first node that does not return 0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ccv_b200 import abi, nnc, resnet50

nnc.init()
stream = nnc.Stream(0)
net = resnet50.Net(4, image=96, classes=10, seed=7, dtype=abi.CCV_16BF)
names = dict((getattr(abi, n), n) for n in dir(abi) if n.startswith("CCV_NNC_") and isinstance(getattr(abi, n), int))
for tag, nodes in (("fwd", net.fwd), ("bwd", net.bwd), ("opt", net.opt)):
    for i, (cmd, hint, flags, ins, outs) in enumerate(nodes):
        st = nnc.cmd_exec(cmd, hint, flags, ins, outs, stream)
        stream.wait()
        if st != 0:
            print(tag, i, names.get(cmd.cmd, hex(cmd.cmd)), "returned", st, nnc.lib().ccv_nnc_sm100_last_error())
            print("  inputs", [(t.dims, hex(t.params.datatype)) if t is not None else None for t in ins])
            print("  outputs", [(t.dims, hex(t.params.datatype)) if t is not None else None for t in outs])
            sys.exit(1)
print("unfused: all nodes returned 0; loss", net.loss.download())
g = nnc.Graph()
for n in net.fwd + net.bwd:
    g.exec_new(*n)
g.fuse()
print("fused run:", g.run(stream))
stream.wait()
