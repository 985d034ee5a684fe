"""CPU: checkpoint interop with the UNMODIFIED reference.  The reference writes a torch archive
(NN::write, nn.cpp:189-202); kami_b200/checkpoint.py turns it into the flat blob / KB20 file this repo loads
and back into an archive the reference's NN::read (nn.cpp:204-222) accepts, generation included."""
import numpy as np
import pytest

import harness as H
import nn_oracle as NO

from kami_b200 import checkpoint as cc

pytestmark = pytest.mark.skipif(H.ref_nn_lib() is None, reason="oracle/_ref not built")


def test_blob_layout_is_the_oracles():
    for F, R in ((64, 2), (256, 20), (128, 0)):
        assert cc.param_order(F, R) == NO.param_order(F, R)


def test_reference_archive_to_blob(tmp_path):
    F, R = 64, 2
    params = NO.init_params(F, R, seed=9)
    ref = H.RefNN(F, R, seed=1)
    ref.set_params(params)
    path = tmp_path / "model.pt"
    ref.write(path)
    got, f, r, gen = cc.archive_to_params(path)
    assert (f, r, gen) == (F, R, 0)
    for name, _ in NO.param_order(F, R):
        assert np.array_equal(got[name], params[name]), name
    kb = tmp_path / "model.kb20"
    cc.write_kb20(kb, NO.pack_blob(got, F, R), F, R, gen)
    blob, f2, r2, g2 = cc.read_kb20(kb)
    assert (f2, r2, g2) == (F, R, 0) and np.array_equal(blob, NO.pack_blob(params, F, R))


def test_blob_to_archive_read_by_reference(tmp_path):
    F, R = 64, 1
    params = NO.init_params(F, R, seed=10)
    path = tmp_path / "from_blob.pt"
    cc.params_to_archive(params, F, R, 7, path)
    ref = H.RefNN(F, R, seed=3)   # different random weights
    ref.read(path)
    assert ref.generation() == 7
    got = ref.get_params()
    for name, _ in NO.param_order(F, R):
        assert np.array_equal(got[name], params[name]), name
    # and the network the reference now runs is the one the blob describes
    obs = np.stack([e.observe() for e in H.sample_positions(4, seed=5)])
    rp, rv = ref.forward_full(obs)
    op, ov = NO.forward(params, obs)
    assert np.abs(rp - op).max() < 1e-5 and np.abs(rv - ov).max() < 1e-5


@pytest.mark.gpu
def test_reference_checkpoint_drives_the_cuda_network_and_back(kb, tmp_path):
    """model.pt written by the UNMODIFIED reference (NN::write) -> kami_b200.checkpoint -> kb_net: NN::infer agrees with the
    reference's own LibTorch network on the same weights (north_star tolerances); and the other way: a KB20 file ->
    torch archive -> the reference's NN::read reproduces the tensors and the generation."""
    F, R = 64, 1
    ref = H.RefNN(F, R, seed=4)
    path = tmp_path / "ref_model.pt"
    ref.write(path)
    params, f, r, gen = cc.archive_to_params(path)
    assert (f, r) == (F, R)
    net = kb.NN(F, R)
    net.load_blob(cc.pack_blob(params, F, R))
    obs = np.stack([e.observe() for e in H.sample_positions(24, seed=6)])
    rp, rv = ref.forward_full(obs)
    pol, val = net.forward_full(obs)
    assert np.abs(val - rv).max() <= 1e-2
    assert float((rp * (np.log(rp + 1e-30) - np.log(pol + 1e-30))).sum(1).max()) <= 1e-3
    kb20 = tmp_path / "model.kb20"
    cc.write_kb20(kb20, cc.pack_blob(params, F, R), F, R, 5)
    blob, f2, r2, g2 = cc.read_kb20(kb20)
    back = tmp_path / "back.pt"
    cc.params_to_archive(cc.unpack_blob(blob, f2, r2), f2, r2, g2, back)
    ref2 = H.RefNN(F, R, seed=99)
    ref2.read(back)
    assert ref2.generation() == 5
    got = ref2.get_params()
    for name, _ in cc.param_order(F, R):
        assert np.array_equal(got[name], params[name]), name
