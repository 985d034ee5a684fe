"""Does the tower pay for a cold instruction cache?  clock64 item totals of k_tower64 (CTA 0) for dense forwards issued back
to back, against forwards that each follow a step of an unrelated tree pool (whose kernels refill the instruction caches)."""
import ctypes as C, sys
sys.path[:0] = [".", "tests", "oracle"]
import numpy as np, kami_b200, bench
from kami_b200 import api

api.init(0)
L = kami_b200.lib()
net = kami_b200.NN(64, 2); net.load_blob(bench.random_blob(64, 2, seed=1))
other = kami_b200.NN(64, 2); other.load_blob(bench.random_blob(64, 2, seed=2))
kw = dict(noise_weight=0.05, selfplay_nodes=1024, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(1024, 1 << 19, api.tree_cfg(seed=1000, **kw))
pool.step(other, 300)
obs = np.random.RandomState(0).rand(1024, 1920).astype(np.float32)
L.kb_net_debug_timestamps(net.h, 1, None, 0, None)

def total():
    ts = (C.c_longlong * 128)(); n = C.c_int()
    L.kb_net_debug_timestamps(net.h, 1, ts, 128, C.byref(n))
    raw = np.array(ts[:60]); k = int(np.count_nonzero(raw))
    return int(raw[k - 1] - raw[0])

net.forward_full(obs)
a = []
for _ in range(6):
    net.forward_full(obs); a.append(total())
b = []
for _ in range(6):
    pool.step(other, 2); L.kb_dev_sync()
    net.forward_full(obs); b.append(total())
print("back to back      :", a)
print("after a pool step :", b)
