import sys, ctypes as C
sys.path.insert(0,'.')
import numpy as np, kami_b200 as kb
from kami_b200 import api
sys.path.insert(0,'/root/repo'); import bench
api.init(0); L=kb.lib()
F,R,B=256,2,1024
tr=kb.Trainer(F,R,B); tr.load_blob(bench.random_blob(F,R,1))
rng=np.random.RandomState(0)
obs=(rng.rand(B,1920)<0.1).astype(np.float32); pi=np.zeros((B,4672),np.float32); pi[np.arange(B),rng.randint(0,4672,B)]=1; z=rng.choice([-1.0,1.0],B).astype(np.float32)
for _ in range(2):
    print(tr.forward_backward(obs,pi,z)); tr.apply_sgd(0.002)
print("ok")
