"""Resident step time as the synthetic games age: throughput and profiled phase times every `chunk` steps from fresh trees
(bench.py's state preparation is chosen from this).  Usage: python tools/steady_state.py [total_steps] [chunk]"""
import ctypes as C
import sys

sys.path[:0] = [".", "tests", "oracle"]
import kami_b200
from kami_b200 import api
from bench import FILTERS, NODE_CAPACITY, RESIDUALS, SELFPLAY_NODES, TREES_PER_GPU, random_blob

total = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
cap = int(sys.argv[3]) if len(sys.argv) > 3 else 0
age = int(sys.argv[4]) if len(sys.argv) > 4 else 0       # steps at a 32-node budget first (bench.py's quick aging)
groups = int(sys.argv[5]) if len(sys.argv) > 5 else 1
api.init(0)
L = kami_b200.lib()
net = kami_b200.NN(FILTERS, RESIDUALS)
net.load_blob(random_blob(FILTERS, RESIDUALS, seed=1))
kw = dict(noise_weight=0.05, selfplay_nodes=SELFPLAY_NODES, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20,
          draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(TREES_PER_GPU, NODE_CAPACITY, api.tree_cfg(seed=1000, **kw))
pool.set_terminal_cap(cap)
pool.set_step_groups(groups)
if age:
    pool.set_selfplay_nodes(32)
    pool.step(net, age)
    pool.set_selfplay_nodes(SELFPLAY_NODES)
    print("aged %d steps at 32 nodes: %s" % (age, {k: pool.stats()[k] for k in ("moves", "games", "skipped_leaves")}), flush=True)
ms = C.c_float()
done = 0
while done < total:
    pool.reset_stats()
    L.kb_dev_sync()
    L.kb_timer_start()
    pool.step(net, chunk)
    L.kb_timer_stop(C.byref(ms))
    st = pool.stats()
    done += chunk
    pool.set_profiling(True)
    pool.step(net, 64)
    ph = pool.phase_ms()
    pool.set_profiling(False)
    done += 64
    print("cap %d groups %d steps %6d: %.1f us/step %.2f M evals/s (skipped %.4f) | moves %5d games %4d scanned/eval %.1f depth %.2f | select %.1f tower %.1f expand %.1f us" % (
        cap, groups, done, ms.value / chunk * 1e3, st["evals"] / ms.value / 1e3, st["skipped_leaves"] / (TREES_PER_GPU * chunk), st["moves"], st["games"], st["children_scanned"] / max(1, st["evals"]),
        st["path_nodes"] / max(1, st["evals"]) - 1, ph["select"] * 1e3, ph["tower"] * 1e3, ph["expand"] * 1e3), flush=True)
