"""CPU: pins oracle/train_oracle.py (one mini-batch of NN::train) against the UNMODIFIED reference
NN::train compiled on LibTorch (oracle/_ref/libkami_ref_nn.so): same weights, same batch, one epoch of
one mini-batch -> every parameter and BatchNorm buffer after the step must agree."""
import numpy as np
import pytest

import harness as H
import nn_oracle as NO
import train_oracle as TO


def _batch(n, seed):
    envs = H.sample_positions(n, seed=seed)
    obs = np.stack([e.observe() for e in envs])
    pi, z = TO.synthetic_targets(n, seed + 1, [e.actions() for e in envs])
    return obs, pi, z


@pytest.mark.skipif(H.ref_nn_lib() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("F,R,n,mlr", [(64, 2, 8, 2), (64, 1, 24, 5)])
def test_train_oracle_matches_reference_nn_train(F, R, n, mlr):
    params = NO.init_params(F, R, seed=5)
    ref = H.RefNN(F, R, seed=1)
    ref.set_params(params)
    obs, pi, z = _batch(n, seed=40)
    assert ref.train(obs, pi, z, mlr, 1, n) == 1
    got = ref.get_params()
    want, loss, grads = TO.train_step(params, obs, pi, z, F, R, mlr / 1000.0)
    assert np.isfinite(loss)
    worst = 0.0
    for name, v in want.items():
        d = float(np.abs(got[name] - v).max())
        worst = max(worst, d)
        assert d <= 2e-5 * max(1.0, float(np.abs(v).max())), (name, d)
    # the step must actually have moved the weights
    assert max(float(np.abs(want[k] - params[k]).max()) for k in want if TO.trainable(k)) > 1e-5
    print("train oracle vs reference NN::train: worst abs diff %.2e, loss %.4f" % (worst, loss))


def test_train_oracle_loss_decreases_on_a_fixed_batch():
    F, R, n = 64, 1, 16
    params = NO.init_params(F, R, seed=3, randomize_bn=False)
    obs, pi, z = _batch(n, seed=7)
    losses = []
    for _ in range(4):
        params, loss, _ = TO.train_step(params, obs, pi, z, F, R, 0.002)
        losses.append(loss)
    assert losses[-1] < losses[0]
