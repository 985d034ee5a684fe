import sys, ctypes as C
sys.path.insert(0,'tests'); sys.path.insert(0,'oracle'); sys.path.insert(0,'.')
import numpy as np, kami_b200 as kb, nn_oracle as NO, train_oracle as TO, harness as H
from kami_b200 import api
api.init(0); L=kb.lib()
F,R,B=256,20,1024
params=NO.init_params(F,R,seed=1,randomize_bn=False)
tr=kb.Trainer(F,R,B); tr.load_blob(NO.pack_blob(params,F,R))
envs=H.sample_positions(64,seed=3)
obs=np.stack([e.observe() for e in envs]); pi,z=TO.synthetic_targets(64,4,[e.actions() for e in envs])
obs=np.tile(obs,(B//64,1)); pi=np.tile(pi,(B//64,1)); z=np.tile(z,B//64)
def dev(a):
    p=C.c_void_p(); api._ck(L.kb_dev_alloc(C.byref(p), a.nbytes)); api._ck(L.kb_dev_upload(p, a.ctypes.data_as(C.c_void_p), a.nbytes)); return p
try:
    od,pd,zd=dev(np.ascontiguousarray(obs,np.float32)),dev(np.ascontiguousarray(pi,np.float32)),dev(np.ascontiguousarray(z,np.float32))
except Exception as e:
    print("upload helper missing:",e); raise
for i in range(2):
    print("warm loss", tr.forward_backward_dev(od,pd,zd,B)); tr.apply_sgd(0.002)
ms=C.c_float(); L.kb_dev_sync(); L.kb_timer_start()
K=5
for i in range(K):
    tr.forward_backward_dev(od,pd,zd,B,want_loss=False); tr.apply_sgd(0.002)
L.kb_timer_stop(C.byref(ms))
print("train step 20x256 batch %d: %.2f ms/step, %.0f samples/s"%(B, ms.value/K, B*K/(ms.value*1e-3)))
# flops: forward tower+heads 3034 MFLOP/pos; backward ~2x
print("approx %.0f TFLOP/s (3x forward flops)"%(3*3034.2e6*B*K/(ms.value*1e-3)/1e12))
