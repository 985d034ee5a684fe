"""Generates the committed golden fixtures from the UNMODIFIED reference compiled into
oracle/_ref (run in the dev container, where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

Fixtures (all small):
  encoding_game.json   the reference's own deterministic golden: test/encoding.cpp's 258-move
                       game (moves + sha256 of its stdout)
  positions.npz        400 sampled positions: compact position bytes, reference planes, legal
                       action lists, terminal flags, static eval, keys
  mcts_known.json      root visit counts / whole-tree digests of reference MCTS runs fed
                       deterministic policy/value streams (noise off)
  nn_f64r2.npz         reference NN (LibTorch, fp32 CPU) outputs for seeded weights and inputs
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import harness as H  # noqa: E402
import nn_oracle as NO  # noqa: E402


def encoding_game():
    # must run first: test/encoding.cpp never calls srand, so it consumes the rand() stream
    # right after the 6344 draws of ncZobristInit
    R = H.ref_core()
    e = H.RefEnv()
    out = ["Initialized neocortex lookup tables", "Starting action test"]
    moves, actions = [], []
    while not e.terminal()[0]:
        acts = e.actions()
        for a in acts:
            assert e.encode(e.decode(int(a))) == a
        a = int(acts[R.ref_rand() % len(acts)])
        moves.append(H.uci(e.decode(a)))
        actions.append(a)
        out.append("Pushing " + moves[-1])
        e.push(a)
    out.append("Done")
    sha = hashlib.sha256(("\n".join(out) + "\n").encode()).hexdigest()
    assert sha == "26545df87423d191c2111f136c378866715ed3485dbc85ee0c76dde959c0e409", sha
    json.dump({"sha256": sha, "moves": moves, "actions": actions, "final": e.terminal()[1:]},
              open(os.path.join(HERE, "encoding_game.json"), "w"))
    print("encoding game:", len(moves), "moves", sha[:12])


def positions(n=400, seed=11):
    rng = np.random.RandomState(seed)
    pos, planes, acts, cnts, term, evals, keys = [], [], [], [], [], [], []
    while len(pos) < n:
        r, o = H.RefEnv(), H.OracleEnv()
        target = int(rng.randint(0, 200))
        for _ in range(target):
            if r.terminal()[0]:
                break
            a = r.actions()
            x = int(a[rng.randint(len(a))])
            r.push(x)
            o.push(x)
        t = r.terminal()
        a = r.actions() if not (t[0] and t[2] <= 3) else np.zeros(0, np.int32)
        pos.append(o.export())  # compact wire form; every field is cross-checked against the reference below
        assert o.key() == r.key() and o.board() == r.board()
        planes.append(r.observe())
        row = np.full(128, -1, np.int32)
        row[:len(a)] = a
        acts.append(row)
        cnts.append(len(a))
        term.append([int(t[0]), t[2], t[1]])
        evals.append(r.eval())
        keys.append(r.key())
    np.savez_compressed(os.path.join(HERE, "positions.npz"), pos=np.stack(pos), planes=np.stack(planes).astype(np.int8),
                        actions=np.stack(acts).astype(np.int16), counts=np.array(cnts, np.int16),
                        terminal=np.array(term, np.float32), eval=np.array(evals, np.int32), key=np.array(keys, np.uint64))
    print("positions:", n, "terminal", int(sum(t[0] for t in term)))


def mcts_stream(seed):
    rng = np.random.RandomState(seed)

    def nxt(vmode):
        if not vmode:
            return np.full(H.PSIZE, 1.0 / H.PSIZE, np.float32), 0.0
        p = rng.rand(H.PSIZE).astype(np.float32)
        return (p / p.sum()).astype(np.float32), float(np.float32(rng.rand() * 2 - 1))
    return nxt


def run_mcts(tree, budget, moves, seed, vmode):
    nxt = mcts_stream(seed)
    rec = []
    for _ in range(moves):
        while tree.n() < budget:
            ok, _obs = tree.select()
            if not ok:
                continue
            p, v = nxt(vmode)
            tree.expand(p, v)
        a, n, w, p = tree.root_children()
        d, c = tree.digest()
        pick = tree.pick(0.0)
        rec.append({"actions": a.tolist(), "visits": n.tolist(), "digest": str(d), "nodes": c, "pick": pick,
                    "root_w": float(tree.root_w())})
        tree.push(pick)
        if tree.env.terminal()[0]:
            break
    return rec


MCTS_CASES = {
    "code_defaults_uniform": (dict(noise_weight=0.0), 1024, 3, 0, False),
    "def_yml_uniform": (dict(noise_weight=0.0, **H.DEF_YML), 1024, 3, 0, False),
    "def_yml_random": (dict(noise_weight=0.0, **H.DEF_YML), 256, 10, 1, True),
    "force_scale_random": (dict(noise_weight=0.0, cpuct=2.5, force_expand_unvisited=1, scale_cpuct_by_actions=1,
                                unvisited_node_value_pct=30, bootstrap_weight=35), 128, 12, 2, True),
}


def mcts_known():
    out = {}
    for name, (cfg, budget, moves, seed, vmode) in MCTS_CASES.items():
        out[name] = run_mcts(H.RefMcts(H.default_cfg(**cfg)), budget, moves, seed, vmode)
        print("mcts", name, out[name][0]["visits"][:6], out[name][-1]["nodes"])
    json.dump(out, open(os.path.join(HERE, "mcts_known.json"), "w"))


def nn_golden():
    for F, R, B in ((64, 2, 12),):
        nn = H.RefNN(F, R, seed=1)
        params = NO.init_params(F, R, seed=7)
        nn.set_params(params)
        envs = H.sample_positions(B, seed=3)
        obs = np.stack([e.observe() for e in envs])
        pol, val = nn.forward_full(obs)
        ip, iv = nn.infer(obs)
        np.savez_compressed(os.path.join(HERE, "nn_f%dr%d.npz" % (F, R)), obs=obs.astype(np.int8), policy=pol, value256=val,
                            infer_value=iv, param_seed=7, filters=F, residuals=R)
        print("nn", F, R, pol.shape, float(val[0, 0]))


if __name__ == "__main__":
    assert H.ref_core() is not None, "build oracle/_ref first: make -C oracle ref"
    encoding_game()
    positions()
    mcts_known()
    nn_golden()
