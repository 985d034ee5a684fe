"""CPU: checkpoint interop with the UNMODIFIED reference.  The reference writes a torch archive
(NN::write, nn.cpp:189-202); oracle/checkpoint_convert.py turns it into the flat blob / KB20 file this repo loads
and back into an archive the reference's NN::read (nn.cpp:204-222) accepts, generation included."""
import numpy as np
import pytest

import harness as H
import nn_oracle as NO

cc = pytest.importorskip("checkpoint_convert")

pytestmark = pytest.mark.skipif(H.ref_nn_lib() is None, reason="oracle/_ref not built")


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
