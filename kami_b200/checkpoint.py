"""Checkpoint interop with the reference -- an offline tool of the product (needs PyTorch for the archive
format only; nothing on the compute path imports it).

Converts between the reference's checkpoint (kami/nn/nn.cpp:189-222: a torch archive of NNModule written with
`mod->save(archive)` plus the IValue "generation") and this repo's formats:

  * the flat fp32 blob of kb_net_load_blob / kb_trainer_load_blob (reference module names, nn_oracle.param_order);
  * the "KB20" file kami::NN::write / read use here (kami/nn/nn.h): int32 {magic, filters, residuals, generation}
    followed by the blob.

    python -m kami_b200.checkpoint to-kb20 model.pt model.kb20
    python -m kami_b200.checkpoint to-torch model.kb20 model.pt

Checked by tests/test_checkpoint_interop.py: both directions through the UNMODIFIED reference (oracle/_ref/libkami_ref_nn.so:
NN::write -> here -> blob, and blob -> here -> NN::read), every tensor compared; on the GPU the converted weights drive
kb_net_infer to the reference network's outputs and a KB20 file written by kami::NN comes back through NN::read.
"""
import struct
import sys

import numpy as np
import torch

KB20_MAGIC = 0x3032424B
NFEATURES = 30


def param_order(filters, residuals):
    """(name, shape) of every fp32 tensor of the flat weight blob kb_net_load_blob / kb_trainer_load_blob consume, in
    order.  Names are the reference's register_module names (nn.cpp:20-23, 45-56); BatchNorm carries weight, bias,
    running_mean, running_var."""
    F = filters
    out = []

    def conv(name, co, ci, k):
        out.append((name + ".weight", (co, ci, k, k)))
        out.append((name + ".bias", (co,)))

    def bn(name, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            out.append((name + "." + s, (c,)))

    conv("conv1", F, NFEATURES, 3)
    bn("batchnorm1", F)
    for i in range(residuals):
        r = "residual%d" % i
        conv(r + ".conv1", F, F, 3)
        bn(r + ".batchnorm1", F)
        conv(r + ".conv2", F, F, 3)
        bn(r + ".batchnorm2", F)
    conv("policyconv", 128, F, 1)
    bn("pbatchnorm", 128)
    conv("policyconv2", 73, 128, 1)
    conv("valueconv", 1, F, 1)
    bn("vbatchnorm", 1)
    out.append(("valuefc.weight", (256, 64)))
    out.append(("valuefc.bias", (256,)))
    return out


def pack_blob(params, filters, residuals):
    return np.concatenate([np.asarray(params[n], np.float32).reshape(-1) for n, _ in param_order(filters, residuals)])


def unpack_blob(blob, filters, residuals):
    params, off = {}, 0
    for name, shape in param_order(filters, residuals):
        n = int(np.prod(shape))
        params[name] = np.asarray(blob[off:off + n], np.float32).reshape(shape)
        off += n
    if off != len(blob):
        raise ValueError("blob has %d floats, filters=%d residuals=%d needs %d" % (len(blob), filters, residuals, off))
    return params


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
    params = {name: tensors[name].to(torch.float32).numpy().reshape(shape).copy() for name, shape in param_order(F, R)}
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
        self.conv1 = torch.nn.Conv2d(NFEATURES, F, 3, padding=1)
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
    for name, _ in param_order(F, R):
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
        write_kb20(argv[3], pack_blob(params, F, R), F, R, gen)
    else:
        blob, F, R, gen = read_kb20(argv[2])
        params_to_archive(unpack_blob(blob, F, R), F, R, gen, argv[3])
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
