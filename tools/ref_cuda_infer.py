"""The UNMODIFIED reference kami::NN::infer on its own GPU path (LibTorch CUDA, test/nncuda.cpp's configuration) at batch B,
timed in a process of its own (bench.py's config-2 leg calls this: LibTorch's CUDA backend and its allocator stay out of
the bench process).  Prints one JSON line.  Usage: python tools/ref_cuda_infer.py obs.npy [reps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import torch  # noqa: E402,F401  (loads the CUDA libraries LibTorch needs from the wheel's own directories)

import harness as H  # noqa: E402

obs = np.load(sys.argv[1])
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
F, R = int(os.environ.get("KB_F", 64)), int(os.environ.get("KB_R", 2))
LC = H.ref_nn_cuda_lib()
if LC is None:
    print(json.dumps({"unavailable": "oracle/_ref/libkami_ref_nn_cuda.so not built"}))
    sys.exit(0)
nn = H.RefNN(F, R, seed=1, force_cpu=False, lib=LC)
if not nn.is_cuda():
    print(json.dumps({"unavailable": "torch::cuda::is_available() is false in the reference build"}))
    sys.exit(0)
for _ in range(5):
    pol, val = nn.infer(obs)
t0 = time.time()
for _ in range(reps):
    nn.infer(obs)
dt = (time.time() - t0) / reps
print(json.dumps({"ms_per_call": dt * 1e3, "pred_per_sec": len(obs) / dt, "batch": len(obs), "policy_row_sum": float(pol[0].sum()),
                  "what": "unmodified kami::NN::infer on LibTorch CUDA (cuDNN, fp32 weights, LibTorch's default TF32 convolutions), "
                          "same GPU, pageable host buffers"}))
