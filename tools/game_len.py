import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np, kami_b200, bench
from kami_b200 import api
api.init(0)
net = kami_b200.NN(64, 1); net.load_blob(bench.random_blob(64, 1, seed=2))
pool = kami_b200.TreePool(32, 1 << 14, api.tree_cfg(noise_weight=0.05, selfplay_nodes=6, seed=9, **kami_b200.DEF_YML))
lens = []
pool.request_game()
for it in range(6000):
    pool.step(net, 64)
    g = pool.take_game()
    if g is not None:
        lens.append(len(g)); pool.request_game()
st = pool.stats()
print("games", st["games"], "moves", st["moves"], "mean plies/game %.1f" % (st["moves"] / max(1, st["games"])))
print("sampled %d games: max %d, >640: %d, p99 %.0f" % (len(lens), max(lens), sum(l > 640 for l in lens), np.percentile(lens, 99)))
