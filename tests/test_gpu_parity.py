"""GPU (-m gpu): the CUDA path, called through the C ABI, against the oracle on the same seeded
inputs and against the committed golden fixtures.  Bit-exact for planes / legality / eval /
visit counts; stated tolerances for the bf16 network."""
import json
import os

import numpy as np
import pytest

import harness as H
import nn_oracle as NO
from golden.make_golden import MCTS_CASES, mcts_stream

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- Env / encoder / legality -------------------------------------------------------------------
def test_env_random_games_bit_exact(kb):
    rng = np.random.RandomState(3)
    for g in range(3):
        d, o = kb.Env(), H.OracleEnv()
        for ply in range(400):
            td, to = d.terminal(), o.terminal()
            assert td == to, (g, ply, td, to)
            assert np.array_equal(d.observe(), o.observe())
            assert np.array_equal(d.position().view(np.uint8).reshape(-1), o.export())
            if to[0]:
                break
            ad, ao = d.actions(), o.actions()
            assert np.array_equal(ad, ao), (g, ply, ad, ao)
            assert d.bootstrap(1600.0) == o.bootstrap(1600.0)
            a = int(ao[rng.randint(len(ao))])
            assert d.decode(a) == o.decode(a) and d.encode(o.decode(a)) == a
            d.push(a)
            o.push(a)
            if ply % 37 == 5:
                d.pop(); o.pop(); d.push(a); o.push(a)


def test_encoding_game_golden(kb):
    """The reference's own deterministic test (test/encoding.cpp) replayed on the device."""
    g = json.load(open(os.path.join(G, "encoding_game.json")))
    e = kb.Env()
    for a, mv in zip(g["actions"], g["moves"]):
        acts = e.actions()
        assert a in acts
        assert H.uci(e.decode(a)) == mv
        for x in acts[:4]:
            assert e.encode(e.decode(int(x))) == x
        e.push(a)
    t = e.terminal()
    assert t[0] and [t[1], t[2]] == g["final"]


def test_batched_positions_golden(kb):
    from kami_b200 import api

    z = np.load(os.path.join(G, "positions.npz"))
    pos = api.as_positions(z["pos"])
    planes = api.encode_planes(pos)
    assert np.array_equal(planes, z["planes"].astype(np.float32).reshape(len(pos), -1))
    acts, cnt = api.legal_actions(pos)
    ev = api.static_eval(pos)
    assert np.array_equal(ev, z["eval"])
    for i in range(len(pos)):
        t = z["terminal"][i]
        if t[0] and t[1] <= 3:
            continue  # draw by rule: the reference never generates moves there
        assert cnt[i] == z["counts"][i], i
        assert np.array_equal(acts[i][:cnt[i]], z["actions"][i][:cnt[i]]), i


def test_batched_positions_vs_oracle_large(kb):
    from kami_b200 import api

    envs = H.sample_positions(1500, seed=21, max_ply=160)
    pos = api.as_positions(np.stack([e.export() for e in envs]))
    planes = api.encode_planes(pos)
    acts, cnt = api.legal_actions(pos)
    ev = api.static_eval(pos)
    picks = np.zeros(len(envs), np.int32)
    for i, e in enumerate(envs):
        assert np.array_equal(planes[i], e.observe())
        oa = e.actions()
        assert cnt[i] == len(oa) and np.array_equal(acts[i][:cnt[i]], oa), i
        assert ev[i] == e.eval()
        picks[i] = oa[i % len(oa)]
    nxt = api.apply_actions(pos, picks)
    for i, e in enumerate(envs):
        e.push(int(picks[i]))
        assert np.array_equal(nxt[i:i + 1].view(np.uint8).reshape(-1), e.export()), i


def _tall_from_obs(obs):
    """numpy restatement of csrc/layout.cuh: fp32 [n][64][30] planes -> bf16 tall image [items][640][64] (as uint16):
    board slot s, square q sits at pixel (1 + 9 s + q // 8) * 10 + 1 + q % 8; the 16-byte chunk with channels 8j..8j+7
    of pixel px is stored at chunk slot j ^ (px & 7); everything else is zero."""
    n = len(obs)
    items = (n + 6) // 7
    tall = np.zeros((items, 640, 64), np.uint16)
    bf = (np.ascontiguousarray(obs, np.float32).reshape(n, 64, 30).view(np.uint32) >> 16).astype(np.uint16)  # exact in bf16
    for b in range(n):
        item, slot = divmod(b, 7)
        for q in range(64):
            px = (1 + 9 * slot + q // 8) * 10 + 1 + q % 8
            line = np.zeros(64, np.uint16)
            line[:30] = bf[b, q]
            for j in range(8):
                k = j ^ (px & 7)
                tall[item, px, 8 * k:8 * k + 8] = line[8 * j:8 * j + 8]
    return tall


@pytest.mark.parametrize("n", [1, 7, 8, 50, 0])
def test_tall_bf16_encoder_bit_exact(kb, n):
    """north_star kernel 1 in the tower's own input layout: every byte of the bf16 tall image equals the REFERENCE's
    planes (golden fixture generated from the compiled reference's Env::observe) rearranged by the layout rule; pads and
    unused channels stay zero (ragged last item too).  n = 0 takes every golden position."""
    from kami_b200 import api

    z = np.load(os.path.join(G, "positions.npz"))
    pos = api.as_positions(z["pos"])
    planes = z["planes"].astype(np.float32).reshape(len(pos), -1)
    if n:
        idx = np.linspace(0, len(pos) - 1, n).astype(int)
        pos, planes = pos[idx], planes[idx]
    got, tail = api.encode_planes_tall(pos)
    assert np.array_equal(got, _tall_from_obs(planes))
    assert not tail.any()


def test_empty_batches(kb):
    from kami_b200 import api

    pos = np.zeros(0, api.POSITION_DTYPE)
    assert api.encode_planes(pos).shape == (0, 1920)
    assert api.legal_actions(pos)[1].shape == (0,)


# ---- MCTS ----------------------------------------------------------------------------------------
def _cfg(kb, **kw):
    from kami_b200 import api

    return api.tree_cfg(**kw)


@pytest.mark.parametrize("name", list(MCTS_CASES))
def test_mcts_known_answers_golden(kb, name):
    """Visit counts and whole-tree digests identical to the reference when both sides are fed the
    same network outputs (noise off)."""
    cfg, budget, moves, seed, vmode = MCTS_CASES[name]
    known = json.load(open(os.path.join(G, "mcts_known.json")))[name]
    t = kb.MCTS(cfg=_cfg(kb, **cfg))
    nxt = mcts_stream(seed)
    for rec in known:
        while t.n() < budget:
            ok, _ = t.select()
            if not ok:
                continue
            p, v = nxt(vmode)
            t.expand(p, v)
        a, n, w, p = t.root_children()
        assert a.tolist() == rec["actions"] and n.tolist() == rec["visits"]
        d, c = t.digest()
        assert (str(d), c) == (rec["digest"], rec["nodes"])
        assert float(t.root_w()) == rec["root_w"]
        pick = t.pick(0.0)
        assert pick == rec["pick"]
        t.push(pick)


def test_mcts_vs_oracle_with_compaction(kb):
    """Small node capacity so the copying collector runs at (almost) every push."""
    cfg = dict(noise_weight=0.0, **H.DEF_YML)
    t = kb.MCTS(cfg=_cfg(kb, **cfg), node_capacity=4096)
    o = H.OracleMcts(H.default_cfg(**cfg))
    rng = np.random.RandomState(9)
    for mv in range(14):
        while o.n() < 48:
            so, oo = o.select()
            sd, od = t.select()
            assert so == sd
            if not so:
                continue
            assert np.array_equal(oo, od), (mv, o.n(), np.nonzero(oo != od)[0][:8], oo[oo != od][:8], od[oo != od][:8])
            p = rng.rand(H.PSIZE).astype(np.float32)
            p /= p.sum()
            v = float(np.float32(rng.rand() * 2 - 1))
            o.expand(p, v)
            t.expand(p, v)
        assert t.n() == o.n() and t.digest() == o.digest()
        assert np.array_equal(t.snapshot(), o.snapshot())
        # temperature pick with an injected uniform: same cumulative walk on both sides
        u = float(rng.rand())
        a = o.pick(1.0, u)
        assert t.pick(1.0, u) == a
        o.push(a)
        t.push(a)
        assert np.array_equal(t.root_position().view(np.uint8).reshape(-1), o.env.export())
        if o.env.terminal()[0]:
            break


def test_leaf_path_replays_to_the_leaf_position(kb):
    """kb_tree_leaf_path: the reference's Env sits at the leaf after select() (mcts.h:252-254); replaying the
    returned actions from the root reaches the leaf whose planes select() returned."""
    cfg = dict(noise_weight=0.0, **H.DEF_YML)
    t = kb.MCTS(cfg=_cfg(kb, **cfg))
    o = H.OracleMcts(H.default_cfg(**cfg))
    rng = np.random.RandomState(2)
    deepest = 0
    for it in range(60):
        so, oo = o.select()
        sd, od = t.select()
        assert so == sd
        if not so:
            continue
        path = t.leaf_path()
        deepest = max(deepest, len(path))
        e = H.OracleEnv()
        for a in path:
            e.push(int(a))
        assert np.array_equal(e.observe(), od) and np.array_equal(e.export(), o.env.export())
        p = rng.rand(H.PSIZE).astype(np.float32)
        p /= p.sum()
        v = float(np.float32(rng.rand() * 2 - 1))
        o.expand(p, v)
        t.expand(p, v)
        assert len(t.leaf_path()) == 0
    assert deepest >= 2


def test_mcts_errors(kb):
    from kami_b200 import KamiError

    t = kb.MCTS(cfg=_cfg(kb, noise_weight=0.0))
    with pytest.raises(KamiError):  # "no children to pick from" (mcts.h:139)
        t.pick(0.0)
    with pytest.raises(KamiError):  # expand without a selected leaf
        t.expand(np.zeros(H.PSIZE, np.float32), 0.0)
    ok, _ = t.select()
    assert ok
    t.expand(np.full(H.PSIZE, 1.0 / H.PSIZE, np.float32), 0.0)
    with pytest.raises(KamiError):  # "no child for action" (mcts.h:129)
        t.push(4671)


def test_pool_batched_select_expand_vs_oracle(kb):
    """kb_pool_select / kb_pool_expand on 96 trees == 96 independent oracle trees."""
    cfg = dict(noise_weight=0.0, **H.DEF_YML)
    n = 96
    pool = kb.TreePool(n, 1 << 15, _cfg(kb, **cfg))
    orc = [H.OracleMcts(H.default_cfg(**cfg)) for _ in range(n)]
    rng = np.random.RandomState(4)
    for it in range(40):
        pool.select()
        leaves = pool.leaf_positions()
        pol = rng.rand(n, H.PSIZE).astype(np.float32)
        pol /= pol.sum(1, keepdims=True)
        val = (rng.rand(n) * 2 - 1).astype(np.float32)
        for i, o in enumerate(orc):
            while not o.select()[0]:
                pass
            assert np.array_equal(leaves[i:i + 1].view(np.uint8).reshape(-1), o.env.export()), (it, i)
            o.expand(pol[i], float(val[i]))
        pool.expand(pol, val)
    for i in (0, 17, n - 1):
        assert pool.tree(i).digest() == orc[i].digest()


def test_pool_selfplay_moves_with_deferred_compaction_vs_oracle(kb):
    """The budget branch of Selfplay::inference_main (selfplay.cpp:136-192) inside k_pool_select: argmax
    pick (alpha < 0.1), push with subtree reuse, reset at game end -- on arenas so small that the
    block-cooperative collector (k_pool_compact, deferred to every 8th select) runs many times.  Every
    leaf must equal the leaf of an oracle tree driven through the same loop with the same NN outputs."""
    nodes, n = 24, 12
    cfg = dict(noise_weight=0.0, **H.DEF_YML)
    pool = kb.TreePool(n, 1 << 12, _cfg(kb, selfplay_nodes=nodes, alpha_initial=0.0, alpha_decay=1.0, alpha_final=0.0,
                                        alpha_cutoff=0, **cfg))
    orc = [H.OracleMcts(H.default_cfg(**cfg)) for _ in range(n)]
    rng = np.random.RandomState(11)
    moves = 0
    for it in range(600):
        pool.select()
        leaves = pool.leaf_positions()
        pol = rng.rand(n, H.PSIZE).astype(np.float32)
        pol /= pol.sum(1, keepdims=True)
        val = (rng.rand(n) * 2 - 1).astype(np.float32)
        for i, o in enumerate(orc):
            while True:
                if o.n() >= nodes:
                    o.push(o.pick(0.0))
                    moves += 1
                    if o.env.terminal()[0]:
                        o.reset()
                    continue
                if o.select()[0]:
                    break
            assert np.array_equal(leaves[i:i + 1].view(np.uint8).reshape(-1), o.env.export()), (it, i)
            o.expand(pol[i], float(val[i]))
        pool.expand(pol, val)
    st = pool.stats()
    assert st["moves"] == moves and moves > 40 * n
    for i in (0, 5, n - 1):
        assert pool.tree(i).digest() == orc[i].digest()


# ---- network ---------------------------------------------------------------------------------------
def _kl(p, q):
    return float((p * (np.log(p + 1e-30) - np.log(q + 1e-30))).sum(1).max())


def _check_net(kb, F, R, B, seed, tol_v=1e-2, tol_kl=1e-3):
    params = NO.init_params(F, R, seed=seed)
    net = kb.NN(F, R)
    net.load_blob(NO.pack_blob(params, F, R))
    envs = H.sample_positions(B, seed=seed + 1)
    obs = np.stack([e.observe() for e in envs])
    pol, val = net.forward_full(obs)
    op, ov = NO.forward(params, obs)
    dv, kl = float(np.abs(val - ov).max()), _kl(op, pol)
    print("net F=%d R=%d B=%d: max|dvalue|=%.3e  KL(ref||new)=%.3e  max|dpolicy|=%.3e" % (F, R, B, dv, kl, np.abs(pol - op).max()))
    assert np.abs(pol.sum(1) - 1).max() < 1e-4
    assert dv <= tol_v and kl <= tol_kl  # BASELINE.json north_star tolerances
    # legal-renormalised distribution (what MCTS::expand consumes)
    for i, e in enumerate(envs[:16]):
        a = e.actions()
        pr, pn = op[i][a] / op[i][a].sum(), pol[i][a] / pol[i][a].sum()
        assert float((pr * (np.log(pr + 1e-30) - np.log(pn + 1e-30))).sum()) <= tol_kl
    return net, params, obs, (pol, val), (op, ov)


def test_conv1_layer_exact_structure(kb):
    """R = 0: the tower output is conv1+BN+ReLU alone -- isolates the tcgen05 tile mapping."""
    F = 64
    params = NO.init_params(F, 0, seed=5)
    net = kb.NN(F, 0)
    os.environ["KB_NO_FUSED_TOWER"] = "1"  # per-layer kernels keep activations in global memory
    try:
        net.load_blob(NO.pack_blob(params, F, 0))
    finally:
        del os.environ["KB_NO_FUSED_TOWER"]
    envs = H.sample_positions(9, seed=8)
    obs = np.stack([e.observe() for e in envs])
    net.forward_full(obs)
    x = np.ascontiguousarray(obs.reshape(-1, 8, 8, 30).transpose(0, 3, 1, 2))
    ref = np.maximum(NO._bn(NO._conv(x, params["conv1.weight"], params["conv1.bias"]), params, "batchnorm1"), 0)
    for b in (0, 6, 7, 8):  # boards in the first and second item
        inp = net.debug_activation(0, b)
        assert np.array_equal(inp[:30].reshape(30, 8, 8), x[b])
        got = net.debug_activation(1, b).reshape(F, 8, 8)
        err = np.abs(got - ref[b]).max()
        assert err < 0.05, (b, err)


def test_net_golden_reference_outputs(kb):
    """Against outputs of the reference's own LibTorch fp32 network (tests/golden)."""
    z = np.load(os.path.join(G, "nn_f64r2.npz"))
    params = NO.init_params(64, 2, seed=int(z["param_seed"]))
    net = kb.NN(64, 2)
    net.load_blob(NO.pack_blob(params, 64, 2))
    obs = z["obs"].astype(np.float32)
    pol, val = net.forward_full(obs)
    assert np.abs(val - z["value256"]).max() <= 1e-2
    assert _kl(z["policy"], pol) <= 1e-3
    ip, iv = net.infer(obs)
    assert np.abs(iv - z["infer_value"]).max() <= 1e-2
    assert np.array_equal(iv, val[0, :len(iv)])  # (Q1) value[i] = vh.flat[i]


@pytest.mark.parametrize("F,R,B", [(64, 2, 256), (256, 2, 64), (128, 1, 20)])
def test_net_vs_oracle(kb, F, R, B):
    _check_net(kb, F, R, B, seed=12)


@pytest.mark.parametrize("R,B", [(0, 9), (3, 40), (6, 300)])
def test_fused_tower_depths_vs_oracle(kb, R, B):
    """The fused 64-filter kernel takes 0..6 residual blocks; the weight ring's position at the start of a layer (and so
    the wrap-around of the straight-line issue routines) depends on the layer count, and 300 boards = 43 items: one CTA
    per item here, two items per CTA on the 2048-board batch of the variants test."""
    _check_net(kb, 64, R, B, seed=14)


def test_net_fused_kernel_matches_per_layer_kernels(kb):
    """filters == 64 runs as ONE fused kernel (k_tower64); the per-layer kernels (k_conv + heads)
    must give the same network."""
    params = NO.init_params(64, 2, seed=21)
    blob = NO.pack_blob(params, 64, 2)
    fused, layered = kb.NN(64, 2), kb.NN(64, 2)
    fused.load_blob(blob)
    os.environ["KB_NO_FUSED_TOWER"] = "1"
    try:
        layered.load_blob(blob)
    finally:
        del os.environ["KB_NO_FUSED_TOWER"]
    obs = np.stack([e.observe() for e in H.sample_positions(300, seed=30)])
    pf, vf = fused.forward_full(obs)
    pl, vl = layered.forward_full(obs)
    assert np.abs(vf - vl).max() < 1e-5 and np.abs(pf - pl).max() < 1e-6


def test_net_batch_edges(kb):
    """Ragged batches: 1 board, a full item (7), one past it (8)."""
    params = NO.init_params(64, 1, seed=3)
    net = kb.NN(64, 1)
    net.load_blob(NO.pack_blob(params, 64, 1))
    envs = H.sample_positions(8, seed=6)
    obs = np.stack([e.observe() for e in envs])
    full = net.forward_full(obs)
    for b in (1, 7, 8):
        pol, val = net.forward_full(obs[:b])
        assert np.array_equal(pol, full[0][:b]) and np.array_equal(val, full[1][:b])


def test_net_nan_guard(kb):
    from kami_b200 import KamiError

    params = NO.init_params(64, 0, seed=3)
    params["valuefc.bias"][:] = np.nan
    net = kb.NN(64, 0)
    net.load_blob(NO.pack_blob(params, 64, 0))
    with pytest.raises(KamiError):  # nn.cpp:176-180
        net.infer(np.zeros((2, 1920), np.float32))


# ---- the whole loop ---------------------------------------------------------------------------------
def test_pool_step_selfplay_smoke(kb):
    """Device-resident loop of selfplay.cpp:113-200; checks invariants the domain offers."""
    F, R, n = 64, 2, 128
    params = NO.init_params(F, R, seed=2)
    net = kb.NN(F, R)
    net.load_blob(NO.pack_blob(params, F, R))
    cfg = _cfg(kb, noise_weight=0.05, selfplay_nodes=24, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5,
               alpha_cutoff=20, seed=7, **H.DEF_YML)
    pool = kb.TreePool(n, 1 << 14, cfg)
    iters = 400
    pool.step(net, iters)
    s = pool.stats()
    assert s["evals"] == n * iters  # exactly one leaf per tree per batch, batch always full
    assert s["moves"] > n * 5 and s["children_created"] > s["evals"] * 5
    assert s["path_nodes"] >= s["evals"]
    obs, pi, z = pool.drain_samples(64)
    if len(z):
        assert np.allclose(pi.sum(1), 1.0, atol=1e-3) and set(np.unique(z)).issubset({-1.0, 0.0, 1.0})
    # every tree still satisfies n(root) == 1 + sum(child visits) (+ terminal visits at the root = 0)
    for i in (0, n // 2, n - 1):
        t = pool.tree(i)
        a, cn, w, p = t.root_children()
        if len(cn):
            assert t.n() == 1 + int(cn.sum())
            assert abs(float(p.sum()) - 1.0) < 1e-3


def test_pool_tiny_budget_long_run_keeps_arena_invariants(kb):
    """Regression: at tiny node budgets a move is played every few steps and the arena reserve that triggers the deferred
    collector (k_pool_compact, every 8th select) used to be smaller than what eight expansions allocate -> an
    intermittent KB_ERR_CAPACITY.  Thousands of near-random games on small arenas must run through and keep the accounting
    invariants."""
    net = kb.NN(64, 1)
    net.load_blob(NO.pack_blob(NO.init_params(64, 1, seed=2), 64, 1))
    for nodes, cap in ((6, 1 << 14), (2, 1 << 12), (3, 1 << 13)):
        n = 64
        pool = kb.TreePool(n, cap, _cfg(kb, noise_weight=0.05, selfplay_nodes=nodes, seed=9, **H.DEF_YML))
        iters = 12000
        for _ in range(iters // 500):
            pool.step(net, 500)
        s = pool.stats()
        assert s["evals"] == n * iters and s["games"] > 20 and s["samples"] <= s["moves"]


def test_pool_full_size_config3_properties(kb):
    """BASELINE config 3 at full size: 1024 concurrent games, options.def.yml (2x64 tower, 1024-node budget per move),
    through properties that do not need the oracle: exactly one evaluation per tree per step, every tree moves when its
    root reaches the budget, per-tree visit accounting (n(root) = 1 + sum of child visits), priors that sum to 1, finished
    games only with z in {-1, draw_value, +1} -- and the whole run is deterministic: a second pool with the same seeds
    reaches bit-identical trees (same digests), noise on, which would expose any race between trees or kernels."""
    F, R, n, nodes = 64, 2, 1024, 1024
    net = kb.NN(F, R)
    net.load_blob(NO.pack_blob(NO.init_params(F, R, seed=1), F, R))
    kw = dict(noise_weight=0.05, selfplay_nodes=nodes, alpha_initial=1.0, alpha_decay=0.95, alpha_final=0.5, alpha_cutoff=20,
              draw_value_pct=50, seed=77, **H.DEF_YML)
    iters = 2 * nodes + 100  # every tree plays at least two moves
    digests = []
    for rep in range(2):
        pool = kb.TreePool(n, 1 << 17, _cfg(kb, **kw))
        pool.step(net, iters)
        s = pool.stats()
        assert s["evals"] == n * iters
        assert s["moves"] >= 2 * n and s["path_nodes"] >= s["evals"] and s["children_created"] > 10 * s["evals"]
        for i in (0, 1, 333, 512, n - 1):
            t = pool.tree(i)
            a, cn, w, p = t.root_children()
            assert len(cn) > 0 and t.n() == 1 + int(cn.sum()) and t.n() < nodes
            assert abs(float(p.sum()) - 1.0) < 1e-3
        digests.append([pool.tree(i).digest() for i in range(0, n, 37)])
        obs, pi, z = pool.drain_samples(256)
        if len(z):
            assert np.allclose(pi.sum(1), 1.0, atol=1e-3) and set(np.unique(z)).issubset({-1.0, 0.0, 1.0})
        del pool
    assert digests[0] == digests[1]


@pytest.mark.parametrize("F,R,groups,vmode", [(64, 1, 1, 0), (64, 1, 4, 1), (64, 1, 3, 1), (128, 1, 4, 1)])
def test_pool_step_matches_hostio_path(kb, F, R, groups, vmode):
    """kb_pool_step (resident) and kb_pool_step_hostio (reference-shaped host round trip) drive
    identical searches when noise is off.  One group = one NN::infer batch over all trees, so the reference's
    value indexing (Q1, value[i] = vh.flat[i]) matches the resident step; with several pipelined groups each
    group is its own NN::infer batch, which is the same search only under value_index_mode 1 (vh[i][0])."""
    n = 32 if groups != 3 else 31  # ragged last group
    params = NO.init_params(F, R, seed=4)
    net = kb.NN(F, R)
    net.load_blob(NO.pack_blob(params, F, R))
    kw = dict(noise_weight=0.0, selfplay_nodes=16, seed=1, value_index_mode=vmode, **H.DEF_YML)
    a, b = kb.TreePool(n, 1 << 14, _cfg(kb, **kw)), kb.TreePool(n, 1 << 14, _cfg(kb, **kw))
    a.set_policy_mode(1)  # dense softmax, the arithmetic of the host round trip
    a.step(net, 60)
    obs = np.zeros((n, 1920), np.float32)
    pol = np.zeros((n, H.PSIZE), np.float32)
    val = np.zeros(n, np.float32)
    b.set_hostio_groups(groups)
    b.step_hostio(net, 60, obs, pol, val)
    for i in range(n):
        assert a.tree(i).digest() == b.tree(i).digest()
    assert a.stats()["moves"] == b.stats()["moves"]
    # the caller's buffers hold the last iteration's rows: every policy row is a distribution over 4672 actions
    assert np.allclose(pol.sum(1), 1.0, atol=1e-3) and np.isfinite(val).all() and (np.abs(obs).sum(1) > 0).all()


def test_hostio_groups_reproduce_per_thread_value_indexing(kb):
    """Reference mode (Q1): with G groups, tree t of group g is expanded with vh.flat[t - t0(g)] of the GROUP's batch,
    exactly what inference thread g of the reference sees from its own NN::infer call (nn.cpp:186).  Checked by
    running group g's trees alone in a pool of their own with the same per-tree seeds (noise off, argmax picks)."""
    F, R, n, G = 64, 1, 16, 2
    params = NO.init_params(F, R, seed=5)
    net = kb.NN(F, R)
    net.load_blob(NO.pack_blob(params, F, R))
    kw = dict(noise_weight=0.0, selfplay_nodes=12, seed=3, alpha_initial=0.0, alpha_final=0.0, **H.DEF_YML)
    obs = np.zeros((n, 1920), np.float32)
    pol = np.zeros((n, H.PSIZE), np.float32)
    val = np.zeros(n, np.float32)
    whole = kb.TreePool(n, 1 << 14, _cfg(kb, **kw))
    whole.set_hostio_groups(G)
    whole.step_hostio(net, 40, obs, pol, val)
    first = kb.TreePool(n // G, 1 << 14, _cfg(kb, **kw))  # trees 0..7 = group 0 (same tree ids, same seeds)
    first.set_hostio_groups(1)
    first.step_hostio(net, 40, obs, pol, val)
    for i in range(n // G):
        assert whole.tree(i).digest() == first.tree(i).digest()


@pytest.mark.parametrize("F,R", [(64, 2), (128, 1)])
def test_pool_step_legal_softmax_equals_dense_softmax_renormalised(kb, F, R):
    """kb_pool_step's default policy path (softmax over the legal moves only, fused behind the policy head:
    north_star kernel 3) against the dense softmax + MCTS::expand renormalisation (nn.cpp:80, mcts.h:273-296):
    same children, priors equal to fp32 rounding, same visit counts over a short search."""
    n = 40
    params = NO.init_params(F, R, seed=6)
    net = kb.NN(F, R)
    net.load_blob(NO.pack_blob(params, F, R))
    kw = dict(noise_weight=0.0, selfplay_nodes=0, seed=1, **H.DEF_YML)
    a, b = kb.TreePool(n, 1 << 14, _cfg(kb, **kw)), kb.TreePool(n, 1 << 14, _cfg(kb, **kw))
    b.set_policy_mode(1)
    a.step(net, 1)
    b.step(net, 1)
    for i in (0, 7, n - 1):
        aa, an, aw, ap = a.tree(i).root_children()
        ba, bn, bw, bp = b.tree(i).root_children()
        assert np.array_equal(aa, ba) and len(aa) == 20
        assert np.abs(ap - bp).max() <= 2e-6 and abs(float(ap.sum()) - 1.0) < 1e-5
    a.step(net, 24)
    b.step(net, 24)
    same = sum(np.array_equal(a.tree(i).root_children()[1], b.tree(i).root_children()[1]) for i in range(n))
    assert same >= n - 2  # a prior that differs in the last bit may flip an exact PUCT tie


def test_net_256_filters_full_size_batch_properties(kb):
    """BASELINE config-3 size (1024+ boards, here 1100 = 158 items: some CTA pairs take two items, the item count
    is even, the last item has one board) through the per-layer tcgen05 path (k_conv2 pairs): a subset of the
    boards against the fp32 oracle, and the size-independent property that the network acts on boards
    independently -- a permuted batch gives exactly the permuted outputs."""
    F, R, B = 256, 2, 1100
    params = NO.init_params(F, R, seed=31)
    net = kb.NN(F, R)
    net.load_blob(NO.pack_blob(params, F, R))
    base = np.stack([e.observe() for e in H.sample_positions(100, seed=32)])
    rng = np.random.RandomState(33)
    obs = base[rng.randint(0, 100, size=B)]
    pol, val = net.forward_full(obs)
    assert np.abs(pol.sum(1) - 1).max() < 1e-4
    pick = np.r_[0:6, 511:517, B - 6:B]
    op, ov = NO.forward(params, obs[pick])
    assert np.abs(val[pick] - ov).max() <= 1e-2 and _kl(op, pol[pick]) <= 1e-3
    perm = rng.permutation(B)
    pol2, val2 = net.forward_full(obs[perm])
    assert np.array_equal(pol2, pol[perm]) and np.array_equal(val2, val[perm])
