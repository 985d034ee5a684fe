"""CPU: the oracle restatement against the committed golden fixtures (generated from the
unmodified reference by tests/golden/make_golden.py) and the reference's own deterministic
golden (test/encoding.cpp)."""
import hashlib
import json
import os

import numpy as np

import harness as H
import nn_oracle as NO
from golden.make_golden import MCTS_CASES, run_mcts

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_encoding_game_golden():
    g = json.load(open(os.path.join(G, "encoding_game.json")))
    e = H.OracleEnv()
    out = ["Initialized neocortex lookup tables", "Starting action test"]
    for a, mv in zip(g["actions"], g["moves"]):
        assert not e.terminal()[0]
        acts = e.actions()
        assert a in acts
        for x in acts:  # the property test/encoding.cpp:24-34 checks
            assert e.encode(e.decode(int(x))) == x
        assert H.uci(e.decode(a)) == mv
        out.append("Pushing " + mv)
        e.push(a)
    out.append("Done")
    t = e.terminal()
    assert t[0] and [t[1], t[2]] == g["final"]
    assert hashlib.sha256(("\n".join(out) + "\n").encode()).hexdigest() == g["sha256"]


def _replay_to(pos_bytes):
    return pos_bytes


def test_positions_golden():
    z = np.load(os.path.join(G, "positions.npz"))
    # the fixture stores compact positions; the oracle re-derives everything from a replay, so
    # walk seeded games again exactly like make_golden.positions and compare to the stored rows
    rng = np.random.RandomState(11)
    i = 0
    while i < len(z["pos"]):
        o = H.OracleEnv()
        target = int(rng.randint(0, 200))
        for _ in range(target):
            if o.terminal()[0]:
                break
            a = o.actions()
            o.push(int(a[rng.randint(len(a))]))
        assert np.array_equal(o.export(), z["pos"][i])
        assert o.key() == int(z["key"][i])
        assert np.array_equal(o.observe(), z["planes"][i].astype(np.float32))
        t = o.terminal()
        assert [float(t[0]), float(t[2]), t[1]] == z["terminal"][i].tolist()
        if not (t[0] and t[2] <= 3):
            a = o.actions()
            assert len(a) == z["counts"][i] and np.array_equal(a, z["actions"][i][:len(a)])
        assert o.eval() == z["eval"][i]
        i += 1


def test_mcts_known_answers():
    known = json.load(open(os.path.join(G, "mcts_known.json")))
    for name, (cfg, budget, moves, seed, vmode) in MCTS_CASES.items():
        got = run_mcts(H.OracleMcts(H.default_cfg(**cfg)), budget, moves, seed, vmode)
        assert len(got) == len(known[name])
        for a, b in zip(got, known[name]):
            assert a == b, name
    # SURVEY 8(c) known answers, measured on the reference
    assert known["code_defaults_uniform"][0]["visits"][:4] == [52, 52, 52, 51]
    assert known["def_yml_uniform"][0]["visits"] == [47, 47, 54, 54, 54, 54, 47, 47, 46, 47, 55, 54, 53, 54, 47, 47, 50, 58, 58, 50]


def test_repetition_rule():
    # (Q5) the draw fires at the 5th occurrence.  The start position itself never counts: its key
    # is the bare board key (position.c:30) while every later key also mixes in the castle key,
    # so the first position to occur five times is the one after 1. Nf3 (plies 1,5,9,13,17).
    e = H.OracleEnv()

    def find(u):
        for x in e.actions():
            if H.uci(e.decode(int(x))) == u:
                return int(x)
        raise KeyError(u)
    plies = 0
    for u in ["g1f3", "g8f6", "f3g1", "f6g8"] * 6:
        t = e.terminal()
        if t[0]:
            assert t[2] == 2 and t[1] == 0.0
            break
        e.push(find(u))
        plies += 1
    assert plies == 17 and e.repcount() == 4


def test_nn_oracle_golden():
    z = np.load(os.path.join(G, "nn_f64r2.npz"))
    params = NO.init_params(int(z["filters"]), int(z["residuals"]), seed=int(z["param_seed"]))
    pol, val = NO.forward(params, z["obs"].astype(np.float32))
    assert np.abs(pol - z["policy"]).max() < 1e-6
    assert np.abs(val - z["value256"]).max() < 1e-5
    _, iv = NO.infer(params, z["obs"].astype(np.float32))
    assert np.abs(iv - z["infer_value"]).max() < 1e-5
    # (Q1) infer's value is the flat head of the [B,256] tensor: all from position 0 for B <= 256
    assert np.abs(iv - val[0, :len(iv)]).max() == 0
    blob = NO.pack_blob(params, 64, 2)
    assert blob.size == 200396 + 2 * (64 * 5 + 128 + 1)  # parameters + BN running stats
