"""CPU: differential test of the oracle restatement against the unmodified reference compiled
into oracle/_ref (skipped where /root/reference was never built, e.g. a bare checkout)."""
import random

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.skipif(H.ref_core() is None, reason="oracle/_ref not built (make -C oracle ref)")


def test_zobrist_keys_match_reference():
    L, R = H.oracle(), H.ref_core()
    for sq in range(64):
        for p in range(12):
            assert L.ok_zobrist_piece(sq, p) == R.ref_zobrist_piece(sq, p)
    for r in range(16):
        assert L.ok_zobrist_castle(r) == R.ref_zobrist_castle(r)
    for f in range(8):
        assert L.ok_zobrist_ep(f) == R.ref_zobrist_ep(f)
    assert L.ok_zobrist_btm() == R.ref_zobrist_btm()


def test_random_games_bit_exact():
    rng = random.Random(1)
    npos = 0
    for g in range(40):
        a, b = H.OracleEnv(), H.RefEnv()
        while True:
            ta, tb = a.terminal(), b.terminal()
            assert ta == tb, b.fen()
            assert a.key() == b.key() and a.hmc() == b.hmc() and a.castle() == b.castle() and a.ep() == b.ep()
            assert a.board() == b.board() and a.repcount() == b.repcount()
            assert a.eval() == b.eval(), b.fen()
            assert a.bootstrap(1600.0) == b.bootstrap(1600.0)
            assert np.array_equal(a.observe(), b.observe())
            if ta[0]:
                break
            assert a.check() == b.check()
            xa, xb = a.actions(), b.actions()
            assert np.array_equal(xa, xb), b.fen()
            for x in xa:
                m = a.decode(int(x))
                assert m == b.decode(int(x)) and a.encode(m) == x == b.encode(m)
            npos += 1
            x = int(xa[rng.randrange(len(xa))])
            a.push(x)
            b.push(x)
            if rng.random() < 0.05:
                a.pop(); b.pop(); a.push(x); b.push(x)
    assert npos > 5000


@pytest.mark.parametrize("cfg,budget,moves,vmode", [
    (dict(noise_weight=0.0), 256, 3, False),
    (dict(noise_weight=0.0, **H.DEF_YML), 200, 8, True),
    (dict(noise_weight=0.0, cpuct=2.5, force_expand_unvisited=1, scale_cpuct_by_actions=1, unvisited_node_value_pct=30,
          bootstrap_weight=35), 96, 8, True),
])
def test_mcts_whole_tree_digest(cfg, budget, moves, vmode):
    rng = np.random.RandomState(5)
    o, r = H.OracleMcts(H.default_cfg(**cfg)), H.RefMcts(H.default_cfg(**cfg))
    for _ in range(moves):
        while o.n() < budget:
            so, oo = o.select()
            sr, orr = r.select()
            assert so == sr
            if not so:
                continue
            assert np.array_equal(oo, orr)
            if vmode:
                p = rng.rand(H.PSIZE).astype(np.float32)
                p /= p.sum()
                v = float(rng.rand() * 2 - 1)
            else:
                p, v = np.full(H.PSIZE, 1.0 / H.PSIZE, np.float32), 0.0
            o.expand(p, v)
            r.expand(p, v)
        assert o.n() == r.n() and o.root_w() == r.root_w()
        assert o.digest() == r.digest()
        assert np.array_equal(o.snapshot(), r.snapshot())
        a = o.pick(0.0)
        assert a == r.pick(0.0)
        o.push(a)
        r.push(a)
        if o.env.terminal()[0]:
            break


@pytest.mark.skipif(H.ref_nn_lib() is None, reason="oracle/_ref NN not built")
def test_nn_oracle_vs_libtorch():
    import nn_oracle as NO
    for F, R in ((64, 2), (256, 1)):
        nn = H.RefNN(F, R, seed=1)
        assert set(n for n, _ in NO.param_order(F, R)) == set(nn.get_params().keys())
        q = NO.init_params(F, R, seed=9)
        nn.set_params(q)
        obs = np.stack([e.observe() for e in H.sample_positions(4, seed=2)])
        rp, rv = nn.forward_full(obs)
        op, ov = NO.forward(q, obs)
        assert np.abs(rp - op).max() < 1e-6 and np.abs(rv - ov).max() < 1e-5
        ip, iv = nn.infer(obs)
        jp, jv = NO.infer(q, obs)
        assert np.abs(iv - jv).max() < 1e-5
