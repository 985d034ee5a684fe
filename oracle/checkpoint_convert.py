"""TEST INFRASTRUCTURE / OFFLINE TOOL -- not on the product path (needs PyTorch).

Converts between the reference's checkpoint (kami/nn/nn.cpp:189-222: a torch archive of NNModule written with
`mod->save(archive)` plus the IValue "generation") and this repo's formats:

  * the flat fp32 blob of kb_net_load_blob / kb_trainer_load_blob (reference module names, nn_oracle.param_order);
  * the "KB20" file kami::NN::write / read use here (kami/nn/nn.h): int32 {magic, filters, residuals, generation}
    followed by the blob.

    python oracle/checkpoint_convert.py to-kb20 model.pt model.kb20
    python oracle/checkpoint_convert.py to-torch model.kb20 model.pt

PARITY PINNED: tests/test_checkpoint_interop.py round-trips both directions through the UNMODIFIED reference
(oracle/_ref/libkami_ref_nn.so: NN::write -> here -> blob, and blob -> here -> NN::read) and compares every tensor.
"""
import struct
import sys

import numpy as np
import torch

import nn_oracle as NO

KB20_MAGIC = 0x3032424B


def _infer_shape(names_to_tensors):
    F = int(names_to_tensors["conv1.weight"].shape[0])
    R = 0
    while "residual%d.conv1.weight" % R in names_to_tensors:
        R += 1
    return F, R


def archive_to_params(path):
    """torch archive written by the reference's NN::write -> (params dict of numpy fp32, filters, residuals, generation)."""
    m = torch.jit.load(str(path), map_location="cpu")
    tensors = {k: v.detach() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    F, R = _infer_shape(tensors)
    params = {name: tensors[name].to(torch.float32).numpy().reshape(shape).copy() for name, shape in NO.param_order(F, R)}
    generation = int(getattr(m, "generation"))
    return params, F, R, generation


class _Residual(torch.nn.Module):
    def __init__(self, F):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(F, F, 3, padding=1)
        self.conv2 = torch.nn.Conv2d(F, F, 3, padding=1)
        self.batchnorm1 = torch.nn.BatchNorm2d(F)
        self.batchnorm2 = torch.nn.BatchNorm2d(F)

    def forward(self, x):
        skip = x
        x = torch.relu(self.batchnorm1(self.conv1(x)))
        return skip + torch.relu(self.batchnorm2(self.conv2(x)))


class _Module(torch.nn.Module):
    """Same registered names as the reference's NNModule (nn.cpp:36-57)."""

    generation: int

    def __init__(self, F, R, generation):
        super().__init__()
        self.batchnorm1 = torch.nn.BatchNorm2d(F)
        self.vbatchnorm = torch.nn.BatchNorm2d(1)
        self.pbatchnorm = torch.nn.BatchNorm2d(128)
        self.conv1 = torch.nn.Conv2d(NO.NFEATURES, F, 3, padding=1)
        self.valueconv = torch.nn.Conv2d(F, 1, 1)
        self.policyconv = torch.nn.Conv2d(F, 128, 1)
        self.policyconv2 = torch.nn.Conv2d(128, 73, 1)
        self.valuefc = torch.nn.Linear(64, 256)
        for i in range(R):
            setattr(self, "residual%d" % i, _Residual(F))
        self.generation = generation
        self.R = R

    def forward(self, x):
        return x


def params_to_archive(params, F, R, generation, path):
    """params dict -> a torch archive the reference's NN::read (InputArchive + mod->load) accepts."""
    m = _Module(F, R, int(generation))
    state = m.state_dict()
    for name, _ in NO.param_order(F, R):
        state[name].copy_(torch.from_numpy(np.asarray(params[name], np.float32)).reshape(state[name].shape))
    m.eval()
    torch.jit.script(m).save(str(path))


def write_kb20(path, blob, F, R, generation):
    with open(path, "wb") as f:
        f.write(struct.pack("<4i", KB20_MAGIC, F, R, int(generation)))
        f.write(np.ascontiguousarray(blob, np.float32).tobytes())


def read_kb20(path):
    with open(path, "rb") as f:
        magic, F, R, gen = struct.unpack("<4i", f.read(16))
        if magic != KB20_MAGIC:
            raise ValueError("not a KB20 checkpoint")
        blob = np.frombuffer(f.read(), np.float32).copy()
    return blob, F, R, gen


def main(argv):
    if len(argv) != 4 or argv[1] not in ("to-kb20", "to-torch"):
        print(__doc__)
        return 2
    if argv[1] == "to-kb20":
        params, F, R, gen = archive_to_params(argv[2])
        write_kb20(argv[3], NO.pack_blob(params, F, R), F, R, gen)
    else:
        blob, F, R, gen = read_kb20(argv[2])
        params, off = {}, 0
        for name, shape in NO.param_order(F, R):
            n = int(np.prod(shape))
            params[name] = blob[off:off + n].reshape(shape)
            off += n
        params_to_archive(params, F, R, gen, argv[3])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
