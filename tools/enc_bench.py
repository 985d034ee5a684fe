import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kami_b200, bench
from kami_b200 import api
api.init(0); L = kami_b200.lib()
kw = dict(noise_weight=0.05, selfplay_nodes=64, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(1024, 1 << 16, api.tree_cfg(seed=1000, **kw))
net = kami_b200.NN(64, 2); net.load_blob(bench.random_blob(64, 2, seed=1))
pool.step(net, 2000)
pool.select()
print(json.dumps(bench.encoder_leg(api, L, pool, 6452.5, "measured"), indent=1))
