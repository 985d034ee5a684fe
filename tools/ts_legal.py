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
acc=None; macc=np.zeros(32); reps=0
for rep in range(20):
    pool.step(net, 1)
    ts=(C.c_longlong*128)(); n=C.c_int()
    L.kb_net_debug_timestamps(net.h, 1, ts, 128, C.byref(n))
    raw=np.array(ts[:128]); nst=int(np.count_nonzero(raw[:60])); t=raw[:nst]; d=np.diff(t)
    if acc is not None and len(acc) != len(d):  # (the gather head stamps once per pass: the count varies)
        acc = None; reps = 0
    acc = d if acc is None else acc+d
    reps = 1 if acc is d else reps + 1
    macc += raw[64:96]
acc=acc/reps; macc/=20
print("legal mode, mean of 20 steps, total cycles", acc.sum())
if len(acc) != len(names):  # legal-move gather head: no policyconv2 MMA
    names = names[:13] + ["until H complete", "park lists, ring wait, bar", "dot units", "bar", "max, exp, store"]
for nm,x in zip(names,acc): print("%-16s %7d"%(nm,x))
print("MMA warp per layer: cycles waiting for weight blocks / cycles from inputs ready to last MMA issued")
for l in range(7): print("  L%d  wait %6d  issue span %6d" % (l, macc[2*l], macc[2*l+1]))
print("last raw stamps (absolute clock):", [int(x) for x in raw[:nst]][-8:])
print("producer issue times of the last three layers' blocks (absolute clock):", [int(x) for x in raw[96:96+27] if x])
ph=pool.phase_ms(); print(ph)
