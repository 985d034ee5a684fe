"""clock64 stamps of CTA 0's first epilogue thread in k_tower64p (KB_TOWER_PIPE=1) inside the pool step: per 3x3 layer
[half-1 accumulators ready, layer written], then policyconv written, logits MMAs complete, softmax start, end."""
import ctypes as C
import os
import sys

os.environ.setdefault("KB_TOWER_PIPE", "1")
sys.path.insert(0, ".")
import numpy as np, kami_b200, bench
from kami_b200 import api

api.init(0)
L = kami_b200.lib()
net = kami_b200.NN(64, 2)
net.load_blob(bench.random_blob(64, 2, seed=1))
kw = dict(noise_weight=0.05, selfplay_nodes=1024, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(1024, 1 << 19, api.tree_cfg(seed=1000, **kw))
pool.step(net, 600)
L.kb_net_debug_timestamps(net.h, 1, None, 0, None)
names = ["zero-fill"] + sum([["L%d until half 1 ready" % l, "L%d rest (epilogue tiles 2,3)" % l] for l in range(5)], []) + [
    "value conv + policyconv epilogue", "wait logits MMAs", "logits epilogue", "softmax"]
acc = None
for rep in range(20):
    pool.step(net, 1)
    ts = (C.c_longlong * 64)()
    n = C.c_int()
    L.kb_net_debug_timestamps(net.h, 1, ts, 64, C.byref(n))
    d = np.diff(np.array(ts[:n.value]))
    acc = d if acc is None else acc + d
acc = acc / 20
print("k_tower64p, mean of 20 steps, total cycles", acc.sum())
for nm, x in zip(names, acc):
    print("%-36s %7d" % (nm, x))
