"""ctypes bindings used by the tests (and by bench.py's reference arm).

Three libraries:
  * oracle/liboracle.so              -- this repo's plain-C restatement (the checker)
  * oracle/_ref/libkami_ref_core.so  -- the unmodified reference Env/MCTS (optional)
  * oracle/_ref/libkami_ref_nn.so    -- the unmodified reference NN on LibTorch (optional)

Nothing in the product package imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PSIZE = 4672
OBSIZE = 1920

c_int_p = C.POINTER(C.c_int)
c_float_p = C.POINTER(C.c_float)


def _fp(a):
    return a.ctypes.data_as(c_float_p)


def _ip(a):
    return a.ctypes.data_as(c_int_p)


class MctsCfg(C.Structure):
    _fields_ = [
        ("cpuct", C.c_float),
        ("force_expand_unvisited", C.c_int),
        ("unvisited_node_value_pct", C.c_int),
        ("bootstrap_weight", C.c_int),
        ("bootstrap_window", C.c_int),
        ("bootstrap_amp_pct", C.c_int),
        ("scale_cpuct_by_actions", C.c_int),
        ("noise_weight", C.c_float),
        ("seed", C.c_uint64),
    ]


def build_oracle():
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "kami_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        L = C.CDLL(build_oracle())
        L.ok_env_new.restype = C.c_void_p
        L.ok_env_turn.restype = C.c_float
        L.ok_env_bootstrap.restype = C.c_float
        L.ok_env_key.restype = C.c_uint64
        L.ok_mcts_new.restype = C.c_void_p
        L.ok_mcts_env.restype = C.c_void_p
        L.ok_mcts_root_w.restype = C.c_float
        L.ok_mcts_digest.restype = C.c_uint64
        for n in ("ok_zobrist_piece", "ok_zobrist_castle", "ok_zobrist_ep", "ok_zobrist_btm"):
            getattr(L, n).restype = C.c_uint64
        L.ok_env_bootstrap.argtypes = [C.c_void_p, C.c_float]
        L.ok_mcts_expand.argtypes = [C.c_void_p, c_float_p, C.c_float, C.c_int]
        L.ok_mcts_pick.argtypes = [C.c_void_p, C.c_float, C.c_double]
        for n in ("ok_env_free", "ok_env_reset", "ok_env_ply", "ok_env_turn", "ok_env_pop", "ok_env_key",
                  "ok_env_hmc", "ok_env_check", "ok_env_repcount", "ok_env_castle", "ok_env_ep", "ok_env_eval",
                  "ok_mcts_free", "ok_mcts_n", "ok_mcts_reset", "ok_mcts_env", "ok_mcts_root_w"):
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("ok_env_encode", "ok_env_decode", "ok_env_push", "ok_env_piece_at", "ok_mcts_push"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        L.ok_env_observe.argtypes = [C.c_void_p, c_float_p]
        L.ok_env_actions.argtypes = [C.c_void_p, c_int_p, C.c_int]
        L.ok_env_terminal.argtypes = [C.c_void_p, c_float_p, c_int_p]
        L.ok_env_export.argtypes = [C.c_void_p, C.c_void_p]
        L.ok_mcts_new.argtypes = [C.POINTER(MctsCfg)]
        L.ok_mcts_default_cfg.argtypes = [C.POINTER(MctsCfg)]
        L.ok_mcts_select.argtypes = [C.c_void_p, c_float_p]
        L.ok_mcts_snapshot.argtypes = [C.c_void_p, c_float_p]
        L.ok_mcts_root_children.argtypes = [C.c_void_p, c_int_p, c_int_p, c_float_p, c_float_p, C.c_int]
        L.ok_mcts_digest.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        L.ok_init()
        _oracle = L
    return _oracle


def ref_core_path():
    return os.path.join(ORACLE_DIR, "_ref", "libkami_ref_core.so")


def ref_nn_path():
    return os.path.join(ORACLE_DIR, "_ref", "libkami_ref_nn.so")


_ref = None


def ref_core():
    """The compiled reference, or None when oracle/_ref has not been built."""
    global _ref
    if _ref is None:
        p = ref_core_path()
        if not os.path.exists(p):
            return None
        L = C.CDLL(p)
        L.ref_env_new.restype = C.c_void_p
        L.ref_env_turn.restype = C.c_float
        L.ref_env_bootstrap.restype = C.c_float
        L.ref_env_key.restype = C.c_uint64
        L.ref_mcts_new.restype = C.c_void_p
        L.ref_mcts_env.restype = C.c_void_p
        L.ref_mcts_root_w.restype = C.c_float
        L.ref_mcts_digest.restype = C.c_uint64
        for n in ("ref_zobrist_piece", "ref_zobrist_castle", "ref_zobrist_ep", "ref_zobrist_btm"):
            getattr(L, n).restype = C.c_uint64
        L.ref_opt_set_int.argtypes = [C.c_char_p, C.c_int]
        L.ref_opt_set_float.argtypes = [C.c_char_p, C.c_float]
        L.ref_env_bootstrap.argtypes = [C.c_void_p, C.c_float]
        L.ref_mcts_expand.argtypes = [C.c_void_p, c_float_p, C.c_float, C.c_int]
        L.ref_mcts_pick.argtypes = [C.c_void_p, C.c_float]
        for n in ("ref_env_free", "ref_env_ply", "ref_env_turn", "ref_env_pop", "ref_env_key", "ref_env_hmc",
                  "ref_env_check", "ref_env_repcount", "ref_env_castle", "ref_env_ep", "ref_env_eval",
                  "ref_mcts_free", "ref_mcts_n", "ref_mcts_reset", "ref_mcts_env", "ref_mcts_root_w"):
            getattr(L, n).argtypes = [C.c_void_p]
        for n in ("ref_env_encode", "ref_env_decode", "ref_env_push", "ref_env_piece_at", "ref_mcts_push"):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        L.ref_env_observe.argtypes = [C.c_void_p, c_float_p]
        L.ref_env_actions.argtypes = [C.c_void_p, c_int_p, C.c_int]
        L.ref_env_terminal.argtypes = [C.c_void_p, c_float_p]
        L.ref_env_terminal_str.argtypes = [C.c_void_p, c_float_p, C.c_char_p, C.c_int]
        L.ref_env_fen.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.ref_mcts_select.argtypes = [C.c_void_p, c_float_p]
        L.ref_mcts_snapshot.argtypes = [C.c_void_p, c_float_p]
        L.ref_mcts_root_children.argtypes = [C.c_void_p, c_int_p, c_int_p, c_float_p, c_float_p, C.c_int]
        L.ref_mcts_digest.argtypes = [C.c_void_p, C.POINTER(C.c_long)]
        _ref = L
    return _ref


class _EnvBase:
    """Uniform python face over the oracle env and the reference env."""

    prefix = ""
    L = None

    def _f(self, name):
        return getattr(self.L, self.prefix + name)

    def ply(self):
        return self._f("env_ply")(self.h)

    def turn(self):
        return self._f("env_turn")(self.h)

    def push(self, a):
        self._f("env_push")(self.h, int(a))

    def pop(self):
        self._f("env_pop")(self.h)

    def actions(self):
        buf = np.zeros(256, np.int32)
        n = self._f("env_actions")(self.h, _ip(buf), 256)
        return buf[:n].copy()

    def observe(self):
        o = np.zeros(OBSIZE, np.float32)
        self._f("env_observe")(self.h, _fp(o))
        return o

    def encode(self, mv):
        return self._f("env_encode")(self.h, int(mv))

    def decode(self, a):
        return self._f("env_decode")(self.h, int(a))

    def bootstrap(self, window):
        return self._f("env_bootstrap")(self.h, float(window))

    def key(self):
        return self._f("env_key")(self.h)

    def hmc(self):
        return self._f("env_hmc")(self.h)

    def check(self):
        return self._f("env_check")(self.h)

    def repcount(self):
        return self._f("env_repcount")(self.h)

    def castle(self):
        return self._f("env_castle")(self.h)

    def ep(self):
        return self._f("env_ep")(self.h)

    def eval(self):
        return self._f("env_eval")(self.h)

    def board(self):
        return [self._f("env_piece_at")(self.h, s) for s in range(64)]


class OracleEnv(_EnvBase):
    prefix = "ok_"

    def __init__(self, handle=None):
        self.L = oracle()
        self.own = handle is None
        self.h = self.L.ok_env_new() if handle is None else handle

    def __del__(self):
        if getattr(self, "own", False) and self.h:
            self.L.ok_env_free(self.h)
            self.h = None

    def terminal(self):
        v = C.c_float()
        r = C.c_int()
        t = self.L.ok_env_terminal(self.h, C.byref(v), C.byref(r))
        return bool(t), v.value, r.value

    def export(self):
        buf = np.zeros(80, np.uint8)
        self.L.ok_env_export(self.h, buf.ctypes.data_as(C.c_void_p))
        return buf


class RefEnv(_EnvBase):
    prefix = "ref_"

    def __init__(self, handle=None):
        self.L = ref_core()
        self.own = handle is None
        self.h = self.L.ref_env_new() if handle is None else handle

    def __del__(self):
        if getattr(self, "own", False) and self.h:
            self.L.ref_env_free(self.h)
            self.h = None

    _REASONS = {"Draw by 50-move rule": 1, "Draw by threefold repetition": 2,
                "Draw by insufficient material": 3, "White is checkmated": 4, "Black is checkmated": 4,
                "White is stalemated": 5, "Black is stalemated": 5}

    def terminal(self):
        v = C.c_float()
        s = C.create_string_buffer(64)
        t = self.L.ref_env_terminal_str(self.h, C.byref(v), s, 64)
        return bool(t), v.value, (self._REASONS[s.value.decode()] if t else 0)

    def fen(self):
        s = C.create_string_buffer(128)
        self.L.ref_env_fen(self.h, s, 128)
        return s.value.decode()


def default_cfg(**kw):
    c = MctsCfg()
    oracle().ok_mcts_default_cfg(C.byref(c))
    for k, v in kw.items():
        setattr(c, k, v)
    return c


# options.def.yml values for the keys the hot path reads
DEF_YML = dict(cpuct=1.5, force_expand_unvisited=0, unvisited_node_value_pct=50, bootstrap_weight=20,
               bootstrap_window=1600, bootstrap_amp_pct=75, scale_cpuct_by_actions=0)


class _MctsBase:
    prefix = ""

    def _f(self, name):
        return getattr(self.L, self.prefix + name)

    def n(self):
        return self._f("mcts_n")(self.h)

    def select(self):
        o = np.zeros(OBSIZE, np.float32)
        r = self._f("mcts_select")(self.h, _fp(o))
        return bool(r), o

    def expand(self, policy, value, disable_bootstrap=False):
        policy = np.ascontiguousarray(policy, np.float32)
        self._f("mcts_expand")(self.h, _fp(policy), float(value), int(disable_bootstrap))

    def push(self, a):
        return self._f("mcts_push")(self.h, int(a))

    def reset(self):
        self._f("mcts_reset")(self.h)

    def snapshot(self):
        o = np.zeros(PSIZE, np.float32)
        self._f("mcts_snapshot")(self.h, _fp(o))
        return o

    def root_children(self):
        a = np.zeros(256, np.int32)
        n = np.zeros(256, np.int32)
        w = np.zeros(256, np.float32)
        p = np.zeros(256, np.float32)
        k = self._f("mcts_root_children")(self.h, _ip(a), _ip(n), _fp(w), _fp(p), 256)
        return a[:k].copy(), n[:k].copy(), w[:k].copy(), p[:k].copy()

    def root_w(self):
        return self._f("mcts_root_w")(self.h)

    def digest(self):
        c = C.c_long()
        d = self._f("mcts_digest")(self.h, C.byref(c))
        return d, c.value


class OracleMcts(_MctsBase):
    prefix = "ok_"

    def __init__(self, cfg=None):
        self.L = oracle()
        cfg = cfg or default_cfg()
        self.h = self.L.ok_mcts_new(C.byref(cfg))
        self.env = OracleEnv(self.L.ok_mcts_env(self.h))

    def __del__(self):
        if self.h:
            self.L.ok_mcts_free(self.h)
            self.h = None

    def pick(self, alpha=0.0, u01=0.0):
        return self.L.ok_mcts_pick(self.h, float(alpha), float(u01))


class RefMcts(_MctsBase):
    prefix = "ref_"

    def __init__(self, cfg=None):
        self.L = ref_core()
        cfg = cfg or default_cfg()
        L = self.L
        L.ref_opt_set_float(b"cpuct", cfg.cpuct)
        L.ref_opt_set_int(b"force_expand_unvisited", cfg.force_expand_unvisited)
        L.ref_opt_set_int(b"unvisited_node_value_pct", cfg.unvisited_node_value_pct)
        L.ref_opt_set_int(b"bootstrap_weight", cfg.bootstrap_weight)
        L.ref_opt_set_int(b"bootstrap_window", cfg.bootstrap_window)
        L.ref_opt_set_int(b"bootstrap_amp_pct", cfg.bootstrap_amp_pct)
        L.ref_opt_set_int(b"scale_cpuct_by_actions", cfg.scale_cpuct_by_actions)
        L.ref_opt_set_float(b"mcts_noise_weight", cfg.noise_weight)
        self.h = L.ref_mcts_new()
        self.env = RefEnv(L.ref_mcts_env(self.h))

    def __del__(self):
        if self.h:
            self.L.ref_mcts_free(self.h)
            self.h = None

    def pick(self, alpha=0.0, u01=None):
        # the reference draws rand()/RAND_MAX itself (mcts.h:173)
        return self.L.ref_mcts_pick(self.h, float(alpha))


def uci(mv):
    s, d, p = (mv >> 6) & 63, mv & 63, (mv >> 12) & 15
    r = "abcdefgh"[s % 8] + str(s // 8 + 1) + "abcdefgh"[d % 8] + str(d // 8 + 1)
    if p < 6:
        r += "pnbrqk"[p]
    return r


# ---- reference NN (LibTorch) ---------------------------------------------------------------
_refnn = None


def ref_nn_cuda_lib():
    """The same reference NN shim with LibTorch's CUDA backend linked in (oracle/Makefile): the reference's own GPU path,
    loaded only by bench.py's config-2 leg on the GPU box."""
    global _refnn_cuda
    if _refnn_cuda is None:
        _refnn_cuda = _load_ref_nn(os.path.join(os.path.dirname(ref_nn_path()), "libkami_ref_nn_cuda.so"))
    return _refnn_cuda


_refnn_cuda = None


def ref_nn_lib():
    global _refnn
    if _refnn is None:
        _refnn = _load_ref_nn(ref_nn_path())
    return _refnn


def _load_ref_nn(p):
    if True:
        if not os.path.exists(p):
            return None
        L = C.CDLL(p)
        L.ref_nn_new.restype = C.c_void_p
        L.ref_nn_new.argtypes = [C.c_int, C.c_int, C.c_uint64, C.c_int]
        L.ref_nn_free.argtypes = [C.c_void_p]
        L.ref_nn_is_cuda.argtypes = [C.c_void_p]
        L.ref_nn_infer.argtypes = [C.c_void_p, c_float_p, C.c_int, c_float_p, c_float_p]
        L.ref_nn_forward_full.argtypes = [C.c_void_p, c_float_p, C.c_int, c_float_p, c_float_p]
        L.ref_nn_num_tensors.argtypes = [C.c_void_p]
        L.ref_nn_tensor_info.restype = C.c_long
        L.ref_nn_tensor_info.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_long), c_int_p, c_int_p]
        L.ref_nn_tensor_get.argtypes = [C.c_void_p, C.c_int, c_float_p]
        L.ref_nn_tensor_set.argtypes = [C.c_void_p, C.c_int, c_float_p]
        L.ref_nn_set_threads.argtypes = [C.c_int]
        L.ref_nn_write.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_nn_read.argtypes = [C.c_void_p, C.c_char_p]
        L.ref_nn_generation.argtypes = [C.c_void_p]
        L.ref_nn_train.argtypes = [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int]
        return L


class RefNN:
    """The unmodified reference kami::NN (kami/nn/nn.cpp) on LibTorch."""

    def __init__(self, filters, residuals, seed=1, force_cpu=True, lib=None):
        self.L = lib or ref_nn_lib()
        self.filters, self.residuals = filters, residuals
        self.h = self.L.ref_nn_new(filters, residuals, seed, int(force_cpu))
        self.index = {}
        for i in range(self.L.ref_nn_num_tensors(self.h)):
            name = C.create_string_buffer(128)
            dims = (C.c_long * 4)()
            rank, isint = C.c_int(), C.c_int()
            self.L.ref_nn_tensor_info(self.h, i, name, 128, dims, C.byref(rank), C.byref(isint))
            self.index[name.value.decode()] = (i, tuple(dims[:rank.value]), bool(isint.value))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_nn_free(self.h)
            self.h = None

    def is_cuda(self):
        return bool(self.L.ref_nn_is_cuda(self.h))

    def get_params(self):
        out = {}
        for name, (i, shape, isint) in self.index.items():
            if isint:
                continue
            a = np.zeros(shape, np.float32)
            self.L.ref_nn_tensor_get(self.h, i, _fp(a))
            out[name] = a
        return out

    def set_params(self, params):
        for name, v in params.items():
            i, shape, isint = self.index[name]
            a = np.ascontiguousarray(v, np.float32).reshape(shape)
            self.L.ref_nn_tensor_set(self.h, i, _fp(a))

    def infer(self, obs):
        obs = np.ascontiguousarray(obs, np.float32)
        B = obs.size // OBSIZE
        pol = np.zeros((B, PSIZE), np.float32)
        val = np.zeros(B, np.float32)
        rc = self.L.ref_nn_infer(self.h, _fp(obs), B, _fp(pol), _fp(val))
        if rc != 0:
            raise RuntimeError("reference NN::infer threw")
        return pol, val

    def forward_full(self, obs):
        obs = np.ascontiguousarray(obs, np.float32)
        B = obs.size // OBSIZE
        pol = np.zeros((B, PSIZE), np.float32)
        val = np.zeros((B, 256), np.float32)
        self.L.ref_nn_forward_full(self.h, _fp(obs), B, _fp(pol), _fp(val))
        return pol, val

    def write(self, path):
        if self.L.ref_nn_write(self.h, str(path).encode()) != 0:
            raise RuntimeError("reference NN::write threw")

    def read(self, path):
        if self.L.ref_nn_read(self.h, str(path).encode()) != 0:
            raise RuntimeError("reference NN::read threw")

    def generation(self):
        return self.L.ref_nn_generation(self.h)

    def train(self, obs, obs_p, obs_v, mlr, epochs, batchsize):
        """NN::train (nn.cpp:224-377); returns the new generation."""
        obs = np.ascontiguousarray(obs, np.float32)
        obs_p = np.ascontiguousarray(obs_p, np.float32)
        obs_v = np.ascontiguousarray(obs_v, np.float32)
        g = self.L.ref_nn_train(self.h, len(obs_v), _fp(obs), _fp(obs_p), _fp(obs_v), mlr, epochs, batchsize)
        if g < 0:
            raise RuntimeError("reference NN::train threw")
        return g


def sample_positions(n, seed=0, max_ply=120):
    """Synthetic positions as SURVEY.md 8(d) config 2 describes: seeded random legal games
    sampled at plies uniform in [0, max_ply].  Returns a list of OracleEnv."""
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < n:
        e = OracleEnv()
        target = int(rng.randint(0, max_ply + 1))
        ok = True
        for _ in range(target):
            if e.terminal()[0]:
                ok = False
                break
            a = e.actions()
            e.push(int(a[rng.randint(len(a))]))
        if ok and not e.terminal()[0]:
            out.append(e)
    return out
