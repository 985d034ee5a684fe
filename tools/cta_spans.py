import sys, ctypes as C
sys.path.insert(0,'.')
import numpy as np, kami_b200, bench
from kami_b200 import api
api.init(0); L=kami_b200.lib()
net = kami_b200.NN(64,2); net.load_blob(bench.random_blob(64,2,seed=1))
kw = dict(noise_weight=0.05, selfplay_nodes=1024, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
pool = kami_b200.TreePool(1024, 1<<19, api.tree_cfg(seed=1000, **kw))
pool.set_terminal_cap(2)
pool.set_selfplay_nodes(32); pool.step(net, 6000); pool.set_selfplay_nodes(1024); pool.step(net, 3500)
L.kb_net_debug_timestamps(net.h, 1, None, 0, None)
ms=C.c_float()
for rep in range(4):
    L.kb_dev_sync(); L.kb_timer_start(); pool.step(net, 64); L.kb_timer_stop(C.byref(ms))   # the last launch's stamps survive
    print('   step %.1f us' % (ms.value/64*1e3))
    buf=(C.c_longlong*320)(); n=C.c_int()
    L.kb_net_debug_cta_spans(net.h, buf, 160, C.byref(n))
    a=np.array(buf[:2*147]).reshape(147,2)
    st, en = a[:,0], a[:,1]
    dur = en-st
    print("rep %d: kernel span (max end - min start) %.1f us | CTA duration min %.1f median %.1f max %.1f us | start spread %.1f us | end spread %.1f us" % (
        rep, (en.max()-st.min())/1e3, dur.min()/1e3, np.median(dur)/1e3, dur.max()/1e3, (st.max()-st.min())/1e3, (en.max()-en.min())/1e3))
ph=pool.phase_ms(); print(ph)
