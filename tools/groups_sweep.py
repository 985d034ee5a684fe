"""Resident step time over the number of independent tree groups (kb_pool_set_step_groups), 1024 trees, options.def.yml.
Usage: python tools/groups_sweep.py [groups ...]"""
import ctypes as C
import sys

sys.path[:0] = [".", "tests", "oracle"]
import kami_b200
from kami_b200 import api
from bench import FILTERS, NODE_CAPACITY, RESIDUALS, SELFPLAY_NODES, TREES_PER_GPU, random_blob

api.init(0)
L = kami_b200.lib()
net = kami_b200.NN(FILTERS, RESIDUALS)
net.load_blob(random_blob(FILTERS, RESIDUALS, seed=1))
kw = dict(noise_weight=0.05, selfplay_nodes=SELFPLAY_NODES, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20,
          draw_value_pct=50, **kami_b200.DEF_YML)
for g in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    pool = kami_b200.TreePool(TREES_PER_GPU, NODE_CAPACITY, api.tree_cfg(seed=1000, **kw))
    pool.set_step_groups(g)
    pool.step(net, 1536)
    pool.step(net, 200)
    pool.reset_stats()
    ms = C.c_float()
    best = 1e9
    for rep in range(3):
        L.kb_dev_sync()
        L.kb_timer_start()
        pool.step(net, 1000)
        L.kb_timer_stop(C.byref(ms))
        best = min(best, ms.value)
    print("groups %d: %.2f us/step  %.2f M evals/s (best of 3 x 1000 steps)" % (g, best, TREES_PER_GPU * 1000 / best / 1e3), flush=True)
    del pool
