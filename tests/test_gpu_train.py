"""GPU (-m gpu): the CUDA training step (kb_trainer_*, tcgen05 forward / dgrad / wgrad convolutions in bf16)
against oracle/train_oracle.py (PyTorch fp32 restatement of one NN::train mini-batch, pinned against the
unmodified reference NN::train in tests/test_train_oracle.py).

Two comparisons per gradient tensor:
  * against the oracle with bf16 rounding emulated at the points where the CUDA step stores bf16 (conv
    operands, activations, activation gradients): relative L2 error <= 8 %, cosine >= 0.995 -- the check of the
    kernels.  Typical agreement is 0.05-0.7 %; the slack is for ReLU kinks: run-to-run fp32 summation order
    (atomics) moves pre-activations by a bf16 ulp, and one value-head element crossing zero moves
    valueconv.weight's gradient by ~3 % at 64 boards (measured, bimodal between identical runs);
  * against the fp32 oracle (the reference arithmetic): cosine >= 0.99 and relative L2 error <= 15 % -- what
    bf16 training costs on a random-init network; the error grows towards the input as rounding accumulates.
Loss within 2 % of the fp32 oracle; BatchNorm running statistics within 2e-2 absolute."""
import numpy as np
import pytest

import harness as H
import nn_oracle as NO
import train_oracle as TO

pytestmark = pytest.mark.gpu


def _batch(n, seed):
    envs = H.sample_positions(n, seed=seed)
    obs = np.stack([e.observe() for e in envs])
    pi, z = TO.synthetic_targets(n, seed + 1, [e.actions() for e in envs])
    return obs, pi, z


def _unpack(blob, F, R):
    out, off = {}, 0
    for name, shape in NO.param_order(F, R):
        n = int(np.prod(shape))
        out[name] = blob[off:off + n].reshape(shape)
        off += n
    assert off == blob.size
    return out


@pytest.mark.parametrize("F,R,n", [(128, 1, 37), (256, 2, 64), (64, 2, 30)])
def test_train_step_gradients_match_oracle(kb, F, R, n):
    params = NO.init_params(F, R, seed=8)
    obs, pi, z = _batch(n, seed=50)
    tr = kb.Trainer(F, R, n)
    tr.load_blob(NO.pack_blob(params, F, R))
    loss = tr.forward_backward(obs, pi, z)
    got = _unpack(tr.export_grads(), F, R)
    stats = _unpack(tr.export_blob(), F, R)
    want, wloss, grads = TO.train_step(params, obs, pi, z, F, R, 0.0)
    _, eloss, egrads = TO.train_step(params, obs, pi, z, F, R, 0.0, emulate_bf16=True)
    print("loss gpu %.5f oracle fp32 %.5f bf16-emulated %.5f" % (loss, wloss, eloss))
    bad = []
    for name, g in grads.items():
        d, ge = got[name], egrads[name]
        ng = float(np.linalg.norm(g))
        if ng < 1e-6 * max(1.0, g.size ** 0.5):  # analytically zero (conv bias in front of BatchNorm)
            ok = float(np.abs(d).max()) <= 1e-4
            rel, cos, rele = float(np.abs(d).max()), 1.0, 0.0
        else:
            rel = float(np.linalg.norm(d - g)) / ng
            cos = float((d * g).sum() / (np.linalg.norm(d) * ng + 1e-30))
            rele = float(np.linalg.norm(d - ge)) / float(np.linalg.norm(ge))
            cose = float((d * ge).sum() / (np.linalg.norm(d) * np.linalg.norm(ge) + 1e-30))
            ok = rele <= 0.08 and cose >= 0.995 and rel <= 0.15 and cos >= 0.99
        print("%-34s |g| %.3e  vs bf16-emulated rel %.3e | vs fp32 rel %.3e cos %.5f %s" % (name, ng, rele, rel, cos, "" if ok else "<-- BAD"))
        if not ok:
            bad.append(name)
    for name in want:
        if not TO.trainable(name):
            assert float(np.abs(stats[name] - want[name]).max()) <= 2e-2, name
    assert abs(loss - wloss) <= 0.02 * abs(wloss)
    assert not bad, bad


def test_train_loop_reduces_loss_and_feeds_inference(kb):
    F, R, n = 128, 1, 56
    params = NO.init_params(F, R, seed=2, randomize_bn=False)
    obs, pi, z = _batch(n, seed=60)
    tr = kb.Trainer(F, R, n)
    tr.load_blob(NO.pack_blob(params, F, R))
    losses = []
    for _ in range(6):
        losses.append(tr.forward_backward(obs, pi, z))
        tr.apply_sgd(0.002)
    assert losses[-1] < losses[0] and all(np.isfinite(losses))
    # the oracle walks the same trajectory
    p = params
    ol = []
    for _ in range(6):
        p, l, _ = TO.train_step(p, obs, pi, z, F, R, 0.002)
        ol.append(l)
    assert abs(losses[-1] - ol[-1]) <= 0.03 * abs(ol[-1]), (losses, ol)
    # trained weights (eval mode, running statistics) drive the inference path
    net = kb.NN(F, R)
    net.load_blob(tr.export_blob())
    pol, val = net.forward_full(obs[:8])
    op, ov = NO.forward(p, obs[:8])
    assert np.abs(val - ov).max() <= 3e-2 and np.abs(pol.sum(1) - 1).max() < 1e-4


def test_data_parallel_step_from_one_process(kb):
    """kb_dp_*: a trainer replica per GPU (as many as the box has, at most 2 here), one NCCL all-reduce(sum) of the
    gradient bucket, the same SGD step on every replica -- against two independent single-GPU trainers whose gradients
    are added on the host.  Sums of two fp32 values are order-independent, so with 2 ranks the weights must agree to the
    last bit of the SGD arithmetic; replicas must be bit-identical to each other."""
    F, R, b = 64, 1, 12
    ndev = min(2, kb.device_count())
    params = NO.init_params(F, R, seed=4)
    blob = NO.pack_blob(params, F, R)
    obs, pi, z = _batch(b * ndev, seed=70)
    dp = kb.DataParallelTrainer(list(range(ndev)), F, R, b)
    dp.load_blob(blob)
    lr = 0.002
    loss = dp.step(obs, pi, z, lr)
    got = [dp.export_blob(r) for r in range(ndev)]
    for r in range(1, ndev):
        assert np.array_equal(got[0], got[r]), "replicas diverged"
    total = np.zeros_like(blob)
    want_loss = 0.0
    for r in range(ndev):
        tr = kb.Trainer(F, R, b)
        tr.load_blob(blob)
        want_loss += tr.forward_backward(obs[r * b:(r + 1) * b], pi[r * b:(r + 1) * b], z[r * b:(r + 1) * b])
        total += tr.export_grads()
        stats = tr.export_blob()
    assert abs(loss - want_loss) <= 1e-3 * abs(want_loss)
    # trainable tensors: w - lr * sum of gradients (running statistics are per replica: compare the trainable part only)
    off = 0
    for name, shape in NO.param_order(F, R):
        n = int(np.prod(shape))
        if TO.trainable(name):
            want = blob[off:off + n] - np.float32(lr) * total[off:off + n]
            assert np.abs(got[0][off:off + n] - want).max() <= 2e-6 * max(1.0, float(np.abs(want).max())), name
        off += n
    # a second step runs (communicator reuse) and moves the weights again
    dp.step(obs, pi, z, lr)
    assert not np.array_equal(dp.export_blob(0), got[0])
