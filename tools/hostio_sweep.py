"""hostio e2e: ms/step over the number of pipelined groups (run on the GPU box)."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kami_b200
from kami_b200 import api
import bench

api.init(0)
L = kami_b200.lib()
net = kami_b200.NN(64, 2)
net.load_blob(bench.random_blob(64, 2, seed=1))
kw = dict(noise_weight=0.05, selfplay_nodes=bench.SELFPLAY_NODES, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5,
          alpha_cutoff=20, draw_value_pct=50, **kami_b200.DEF_YML)
n = 1024
pool = kami_b200.TreePool(n, bench.NODE_CAPACITY, api.tree_cfg(seed=1000, **kw))
pool.step(net, 600)
def pinned(shape):
    nbytes = int(np.prod(shape)) * 4
    p = C.c_void_p()
    api._ck(L.kb_host_alloc_pinned(C.byref(p), nbytes))
    return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(int(np.prod(shape)),)).reshape(shape)
obs_h, pol_h, val_h = pinned((n, 1920)), pinned((n, 4672)), pinned((n,))
for g in ([int(os.environ['SWEEP_GROUPS'])] if 'SWEEP_GROUPS' in os.environ else (1, 2, 3, 4, 6, 8, 4)):
    pool.set_hostio_groups(g)
    pool.step_hostio(net, 10, obs_h, pol_h, val_h)
    ms = C.c_float()
    L.kb_dev_sync(); L.kb_timer_start()
    pool.step_hostio(net, 200, obs_h, pol_h, val_h)
    L.kb_timer_stop(C.byref(ms))
    print("groups %d: %.3f ms/step  %.3f M evals/s" % (g, ms.value / 200, n * 200 / ms.value / 1e3), flush=True)
