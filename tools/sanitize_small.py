"""A small tour of the hot path for compute-sanitizer (memcheck): single-tree protocol, pool step with cap and groups,
host-buffer loops (dense and compact), replay drain, arena rounds with two networks, one training step."""
import sys

sys.path[:0] = [".", "tests", "oracle"]
import numpy as np
import kami_b200
from kami_b200 import api
from bench import random_blob

api.init(0)
F, R = 64, 1
net = kami_b200.NN(F, R)
net.load_blob(random_blob(F, R, seed=1))
net2 = kami_b200.NN(F, R)
net2.load_blob(random_blob(F, R, seed=2))
kw = dict(noise_weight=0.05, selfplay_nodes=6, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20, **kami_b200.DEF_YML)
t = kami_b200.MCTS(cfg=api.tree_cfg(noise_weight=0.0, **kami_b200.DEF_YML), node_capacity=2048)
for _ in range(12):
    ok, obs = t.select()
    if ok:
        pol, val = net.infer(obs[None, :])
        t.expand(pol[0], float(val[0]))
pool = kami_b200.TreePool(40, 1 << 12, api.tree_cfg(seed=3, **kw))
pool.set_terminal_cap(2)
pool.step(net, 120)
pool.set_step_groups(3)
pool.step(net, 60)
pool.set_step_groups(0)
pool.set_split_select(1)
pool.step(net, 60)
pool.set_split_select(0)
n = 40
obs, pol, val = np.zeros((n, 1920), np.float32), np.zeros((n, 4672), np.float32), np.zeros(n, np.float32)
pool.step_hostio(net, 8, obs, pol, val)
leaves, prior = np.zeros(n, api.POSITION_DTYPE), np.zeros((n, 128), np.float32)
pool.step_hostio_compact(net, 8, leaves, prior, val)
rows = 0
while True:
    o, p, z = pool.drain_samples(100)
    if not len(z):
        break
    rows += len(z)
pool.flush_trees()
pool.step(net, 20)
arena = api.Arena(6, 4, 8, [1, -1, 1, -1, 1, -1], api.tree_cfg(noise_weight=0.0, **kami_b200.DEF_YML))
games = 0
for _ in range(300):
    games += len(arena.round(net, net2))
tr = kami_b200.Trainer(F, R, 8)
tr.load_blob(random_blob(F, R, seed=1))
pi = np.zeros((8, 4672), np.float32)
pi[:, :20] = 0.05
loss = tr.forward_backward(obs[:8], pi, np.zeros(8, np.float32))
tr.apply_sgd(0.001)
print("sanitize tour ok: stats", {k: pool.stats()[k] for k in ("evals", "moves", "games", "skipped_leaves")}, "rows", rows, "arena games", games, "loss", loss)
