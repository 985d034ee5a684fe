"""The profiled window of the round-2 ncu captures: bench.py's workload (1024 trees, options.def.yml, terminal cap 2) aged
to steady state exactly like bench.py does, then `steps` iterations of the resident step between kb_profiler_start/stop.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/prof_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'k_tower64|k_pool_expand_select' -c 4 -o prof python tools/prof_step.py"""
import sys

sys.path[:0] = [".", "tests", "oracle"]
import kami_b200
from kami_b200 import api
from bench import (AGE_NODES, AGE_STEPS, FILTERS, NODE_CAPACITY, PREROLL_STEPS, RESIDUALS, SELFPLAY_NODES, TERMINAL_CAP, TREES_PER_GPU,
                   random_blob)

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
api.init(0)
L = kami_b200.lib()
net = kami_b200.NN(FILTERS, RESIDUALS)
net.load_blob(random_blob(FILTERS, RESIDUALS, seed=1))
kw = dict(noise_weight=0.05, selfplay_nodes=SELFPLAY_NODES, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20,
          draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(TREES_PER_GPU, NODE_CAPACITY, api.tree_cfg(seed=1000, **kw))
pool.set_terminal_cap(TERMINAL_CAP)
pool.set_selfplay_nodes(AGE_NODES)
pool.step(net, AGE_STEPS)
pool.set_selfplay_nodes(SELFPLAY_NODES)
pool.step(net, PREROLL_STEPS)
L.kb_dev_sync()
L.kb_profiler_start()
pool.step(net, steps)
L.kb_dev_sync()
L.kb_profiler_stop()
print("profiled %d steps: %s" % (steps, {k: pool.stats()[k] for k in ("evals", "moves", "skipped_leaves")}))
