import sys, ctypes as C
sys.path.insert(0,'.')
import numpy as np, kami_b200, bench
from kami_b200 import api
api.init(0); L=kami_b200.lib()
net = kami_b200.NN(64,2); net.load_blob(bench.random_blob(64,2,seed=1))
kw = dict(noise_weight=0.05, selfplay_nodes=1024, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(1024, 1<<19, api.tree_cfg(seed=1000, **kw))
pool.step(net, 600)
L.kb_net_debug_timestamps(net.h, 1, None, 0, None)
names=["zero-fill"]+sum([["L%d wait acc"%l,"L%d epilogue"%l] for l in range(7)],[])+["value+bar","softmax"]
acc=None; macc=np.zeros(32)
for rep in range(20):
    pool.step(net, 1)
    ts=(C.c_longlong*128)(); n=C.c_int()
    L.kb_net_debug_timestamps(net.h, 1, ts, 128, C.byref(n))
    raw=np.array(ts[:128]); nst=int(np.count_nonzero(raw[:60])); t=raw[:nst]; d=np.diff(t)
    acc = d if acc is None else acc+d
    macc += raw[64:96]
acc=acc/20; macc/=20
print("legal mode, mean of 20 steps, total cycles", acc.sum())
for nm,x in zip(names,acc): print("%-16s %7d"%(nm,x))
print("MMA warp per layer: cycles waiting for weight blocks / cycles from inputs ready to last MMA issued")
for l in range(7): print("  L%d  wait %6d  issue span %6d" % (l, macc[2*l], macc[2*l+1]))
ph=pool.phase_ms(); print(ph)
