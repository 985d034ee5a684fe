import sys, ctypes as C
sys.path.insert(0,'.')
import numpy as np, kami_b200, bench
from kami_b200 import api
api.init(0); L=kami_b200.lib()
net = kami_b200.NN(64,2); net.load_blob(bench.random_blob(64,2,1))
kw = dict(noise_weight=0.05, selfplay_nodes=1024, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(1024, 1<<19, api.tree_cfg(seed=1000, **kw))
pool.set_terminal_cap(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
pool.step(net, int(sys.argv[1]) if len(sys.argv) > 1 else 1536)
L.kb_pool_debug_select_profile(pool.h, 1, None, 0)
buf=np.zeros((1024,8),np.int64)
N=80
rows=[]
for i in range(N):
    pool.step(net,1)
    L.kb_pool_debug_select_profile(pool.h, 1, buf.ctypes.data_as(C.c_void_p), 1024)
    t=buf[:,0]
    j=int(t.argmax())
    rows.append((np.median(t), np.percentile(t,90), np.percentile(t,99), t.max(), buf[j,2], buf[j,4], buf[j,6], buf[j,7], buf[j,1], buf[j,3], buf[j,5]))
r=np.array(rows,dtype=float)
srt=np.argsort(-r[:,3])[:8]
for j in srt: print("  slow step: max %.0f cycles  n_sel %d n_move %d depth %d nact %d sel_cyc %.0f move_cyc %.0f"%(r[j,3], r[j,4], -r[j,5] if r[j,5]<0 else r[j,5], r[j,6], r[j,7], r[j,8], abs(r[j,9])))
print("per-step stats over %d steps (cycles): median %.0f p90 %.0f p99 %.0f max %.0f"%(N, *r[:,:4].mean(0)))
print("slowest tree of each step: mean n_sel %.2f n_move %.2f depth %.1f nact %.1f sel_cyc %.0f move_cyc %.0f enc %.0f"%tuple(r[:,4:].mean(0)))
print("steps where the slowest tree made a move: %d, absorbed terminals (n_sel>1): %d"%((r[:,5]>0).sum(), (r[:,4]>1).sum()))
print("all trees last step: mean n_sel %.3f, frac with move %.4f, mean depth %.2f, mean nact %.1f"%(buf[:,2].mean(), (buf[:,4]>0).mean(), buf[:,6].mean(), buf[:,7].mean()))
ok = buf[:,3] > 0
print("descent cycles median %.0f mean %.0f | leaf (terminal test + legal actions) median %.0f mean %.0f | encode median %.0f | total median %.0f"%(np.median(buf[ok,3]), buf[ok,3].mean(), np.median(buf[ok,4]), buf[ok,4].mean(), np.median(buf[ok,5]), np.median(buf[ok,0])))
for d in range(0,8):
    mk = ok & (buf[:,6]==d)
    if mk.sum()>5: print("depth %d: n %d descent %.0f leaf %.0f nact %.1f"%(d, mk.sum(), buf[mk,3].mean(), buf[mk,4].mean(), buf[mk,7].mean()))
