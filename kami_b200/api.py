"""ctypes binding of include/kami_b200.h plus Python mirrors of the reference's classes
(kami::Env env.h:41, kami::MCTS mcts.h:66, kami::NN nn.h:40) with the same method names,
argument meaning and error behaviour, so the parity tests read like the reference's tests."""
import ctypes as C
import os
import subprocess

import numpy as np

PSIZE = 4672
OBSIZE = 1920
NFEATURES = 30
MAX_ACTIONS = 128
VALUE_WIDTH = 256

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# options.def.yml values of the keys the hot path reads
DEF_YML = dict(cpuct=1.5, force_expand_unvisited=0, unvisited_node_value_pct=50, bootstrap_weight=20,
               bootstrap_window=1600, bootstrap_amp_pct=75, scale_cpuct_by_actions=0)


class KamiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kami_b200 error %d: %s" % (code, msg))
        self.code = code


class Position(C.Structure):
    _fields_ = [("pieces", C.c_uint64 * 6), ("white", C.c_uint64), ("board_key", C.c_uint64), ("key", C.c_uint64),
                ("ctm", C.c_uint8), ("castle", C.c_uint8), ("ep", C.c_uint8), ("hmc", C.c_uint8),
                ("ply", C.c_uint16), ("check", C.c_uint8), ("pad", C.c_uint8)]


POSITION_DTYPE = np.dtype([("pieces", "<u8", 6), ("white", "<u8"), ("board_key", "<u8"), ("key", "<u8"),
                           ("ctm", "u1"), ("castle", "u1"), ("ep", "u1"), ("hmc", "u1"), ("ply", "<u2"),
                           ("check", "u1"), ("pad", "u1")])
assert POSITION_DTYPE.itemsize == 80 and C.sizeof(Position) == 80


class TreeCfg(C.Structure):
    _fields_ = [("cpuct", C.c_float), ("force_expand_unvisited", C.c_int), ("unvisited_node_value_pct", C.c_int),
                ("bootstrap_weight", C.c_int), ("bootstrap_window", C.c_int), ("bootstrap_amp_pct", C.c_int),
                ("scale_cpuct_by_actions", C.c_int), ("noise_weight", C.c_float), ("seed", C.c_uint64),
                ("selfplay_nodes", C.c_int), ("alpha_initial", C.c_float), ("alpha_decay", C.c_float),
                ("alpha_final", C.c_float), ("alpha_cutoff", C.c_int), ("draw_value_pct", C.c_int),
                ("value_index_mode", C.c_int)]


class PoolStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("evals", "moves", "games", "terminal_visits", "children_scanned",
                                          "path_nodes", "children_created", "samples", "nodes_in_use",
                                          "kernel_launches", "skipped_leaves")]


class ArenaGame(C.Structure):
    _fields_ = [("tree", C.c_int32), ("result", C.c_float), ("colour", C.c_int32)]


class PhaseMs(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("select", "encode", "tower", "heads", "expand", "total")]


def lib_path():
    return os.path.join(_HERE, "libkami_b200.so")


def build(force=False):
    """Compile libkami_b200.so in-tree with nvcc for sm_100a (kami_b200/csrc/Makefile)."""
    if force:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc")], stdout=subprocess.DEVNULL)
    return lib_path()


# name -> (restype, argtypes); every symbol include/kami_b200.h declares
_P = C.c_void_p
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_SIGS = {
    "kb_init": (C.c_int, [C.c_int]),
    "kb_last_error": (C.c_char_p, []),
    "kb_device_count": (C.c_int, []),
    "kb_device_name": (C.c_int, [C.c_char_p, C.c_int]),
    "kb_sm_count": (C.c_int, []),
    "kb_env_create": (C.c_int, [C.POINTER(_P)]),
    "kb_env_destroy": (C.c_int, [_P]),
    "kb_env_reset": (C.c_int, [_P]),
    "kb_env_ply": (C.c_int, [_P, _i32p]),
    "kb_env_push": (C.c_int, [_P, C.c_int]),
    "kb_env_pop": (C.c_int, [_P]),
    "kb_env_actions": (C.c_int, [_P, _i32p, C.c_int, _i32p]),
    "kb_env_observe": (C.c_int, [_P, _f32p]),
    "kb_env_terminal": (C.c_int, [_P, _i32p, _f32p, _i32p]),
    "kb_env_encode": (C.c_int, [_P, C.c_int, _i32p]),
    "kb_env_decode": (C.c_int, [_P, C.c_int, _i32p]),
    "kb_env_bootstrap": (C.c_int, [_P, C.c_float, _f32p]),
    "kb_env_position": (C.c_int, [_P, _P]),
    "kb_encode_planes": (C.c_int, [_P, C.c_int, _f32p]),
    "kb_legal_actions": (C.c_int, [_P, C.c_int, _i32p, _i32p]),
    "kb_apply_actions": (C.c_int, [_P, C.c_int, _i32p]),
    "kb_static_eval": (C.c_int, [_P, C.c_int, _i32p]),
    "kb_encode_planes_dev": (C.c_int, [_P, C.c_int, _P]),
    "kb_encode_planes_bf16_dev": (C.c_int, [_P, C.c_int, _P]),
    "kb_legal_actions_dev": (C.c_int, [_P, C.c_int, _P, _P]),
    "kb_net_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int]),
    "kb_net_destroy": (C.c_int, [_P]),
    "kb_net_blob_floats": (C.c_size_t, [C.c_int, C.c_int]),
    "kb_net_load_blob": (C.c_int, [_P, _f32p, C.c_size_t]),
    "kb_net_infer": (C.c_int, [_P, _f32p, C.c_int, _f32p, _f32p]),
    "kb_net_forward_full": (C.c_int, [_P, _f32p, C.c_int, _f32p, _f32p]),
    "kb_net_forward_dev": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "kb_net_planes_bytes": (C.c_size_t, [C.c_int]),
    "kb_net_flops": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "kb_trainer_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int]),
    "kb_trainer_destroy": (C.c_int, [_P]),
    "kb_trainer_load_blob": (C.c_int, [_P, C.POINTER(C.c_float), C.c_size_t]),
    "kb_trainer_export_blob": (C.c_int, [_P, C.POINTER(C.c_float), C.c_size_t]),
    "kb_trainer_export_grads": (C.c_int, [_P, C.POINTER(C.c_float), C.c_size_t]),
    "kb_trainer_grad_buffer": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "kb_trainer_forward_backward": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                                              C.POINTER(C.c_float)]),
    "kb_trainer_forward_backward_dev": (C.c_int, [_P, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "kb_trainer_apply_sgd": (C.c_int, [_P, C.c_float, C.c_float]),
    "kb_trainer_param_buffer": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "kb_trainer_stat_ranges": (C.c_int, [_P, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.c_int, _i32p]),
    "kb_trainer_debug_activation": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "kb_net_debug_activation": (C.c_int, [_P, C.c_int, C.c_int, _f32p, _i32p]),
    "kb_net_debug_timestamps": (C.c_int, [_P, C.c_int, C.POINTER(C.c_longlong), C.c_int, _i32p]),
    "kb_net_debug_cta_spans": (C.c_int, [_P, C.POINTER(C.c_longlong), C.c_int, _i32p]),
    "kb_tree_default_cfg": (C.c_int, [C.POINTER(TreeCfg)]),
    "kb_pool_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.POINTER(TreeCfg)]),
    "kb_pool_destroy": (C.c_int, [_P]),
    "kb_pool_size": (C.c_int, [_P]),
    "kb_tree_n": (C.c_int, [_P, C.c_int, _i32p]),
    "kb_tree_select": (C.c_int, [_P, C.c_int, _f32p, _i32p]),
    "kb_tree_expand": (C.c_int, [_P, C.c_int, _f32p, C.c_float, C.c_int]),
    "kb_tree_pick": (C.c_int, [_P, C.c_int, C.c_float, C.c_double, _i32p]),
    "kb_tree_push": (C.c_int, [_P, C.c_int, C.c_int]),
    "kb_tree_reset": (C.c_int, [_P, C.c_int]),
    "kb_tree_snapshot": (C.c_int, [_P, C.c_int, _f32p]),
    "kb_tree_root_children": (C.c_int, [_P, C.c_int, _i32p, _i32p, _f32p, _f32p, C.c_int, _i32p]),
    "kb_tree_root_w": (C.c_int, [_P, C.c_int, _f32p]),
    "kb_tree_leaf_path": (C.c_int, [_P, C.c_int, _i32p, C.c_int, _i32p]),
    "kb_tree_digest": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_int64)]),
    "kb_tree_env": (C.c_int, [_P, C.c_int, _P]),
    "kb_pool_select": (C.c_int, [_P]),
    "kb_pool_leaf_positions": (C.c_int, [_P, _P]),
    "kb_pool_expand": (C.c_int, [_P, _f32p, _f32p, C.c_int]),
    "kb_pool_expand_dev": (C.c_int, [_P, _P, _P, C.c_int]),
    "kb_dp_create": (C.c_int, [C.POINTER(_P), _i32p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "kb_dp_destroy": (C.c_int, [_P]),
    "kb_dp_size": (C.c_int, [_P]),
    "kb_dp_replica": (_P, [_P, C.c_int]),
    "kb_dp_load_blob": (C.c_int, [_P, _f32p, C.c_size_t]),
    "kb_dp_export_blob": (C.c_int, [_P, C.c_int, _f32p, C.c_size_t]),
    "kb_dp_step": (C.c_int, [_P, _f32p, _f32p, _f32p, C.c_int, C.c_float, C.c_float, _f32p]),
    "kb_dp_step_dev": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_float, C.c_float]),
    "kb_dp_allreduce_apply": (C.c_int, [_P, C.c_float, C.c_float]),
    "kb_arena_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.POINTER(TreeCfg), _i32p, C.c_int]),
    "kb_arena_destroy": (C.c_int, [_P]),
    "kb_arena_round": (C.c_int, [_P, _P, _P, C.POINTER(ArenaGame), C.c_int, _i32p]),
    "kb_arena_begin": (C.c_int, [_P, C.POINTER(ArenaGame), C.c_int, _i32p, _f32p, _i32p, _f32p, _i32p]),
    "kb_arena_end": (C.c_int, [_P, _f32p, _f32p, _f32p, _f32p]),
    "kb_arena_pool": (_P, [_P]),
    "kb_arena_colours": (C.c_int, [_P, _i32p, C.c_int]),
    "kb_pool_leaf_actions": (C.c_int, [_P, _i32p, _i32p]),
    "kb_pool_expand_compact": (C.c_int, [_P, _f32p, _f32p, C.c_int]),
    "kb_pool_step": (C.c_int, [_P, _P, C.c_int]),
    "kb_pool_step_hostio_compact": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "kb_pool_set_step_groups": (C.c_int, [_P, C.c_int]),
    "kb_pool_set_profiling": (C.c_int, [_P, C.c_int]),
    "kb_pool_set_terminal_cap": (C.c_int, [_P, C.c_int]),
    "kb_pool_set_selfplay_nodes": (C.c_int, [_P, C.c_int]),
    "kb_pool_set_split_select": (C.c_int, [_P, C.c_int]),
    "kb_pool_flush_trees": (C.c_int, [_P]),
    "kb_current_device": (C.c_int, []),
    "kb_host_register": (C.c_int, [_P, C.c_size_t]),
    "kb_host_unregister": (C.c_int, [_P]),
    "kb_pool_step_hostio": (C.c_int, [_P, _P, C.c_int, _P, _P, _P]),
    "kb_pool_get_stats": (C.c_int, [_P, C.POINTER(PoolStats)]),
    "kb_pool_reset_stats": (C.c_int, [_P]),
    "kb_pool_set_policy_mode": (C.c_int, [_P, C.c_int]),
    "kb_pool_set_hostio_groups": (C.c_int, [_P, C.c_int]),
    "kb_pool_request_game": (C.c_int, [_P]),
    "kb_pool_take_game": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int)]),
    "kb_pool_drain_samples": (C.c_int, [_P, C.c_int, _f32p, _f32p, _f32p, _i32p]),
    "kb_pool_last_phase_ms": (C.c_int, [_P, C.POINTER(PhaseMs)]),
    "kb_pool_debug_select_profile": (C.c_int, [_P, C.c_int, C.c_void_p, C.c_int]),
    "kb_dev_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "kb_dev_free": (C.c_int, [_P]),
    "kb_dev_upload": (C.c_int, [_P, _P, C.c_size_t]),
    "kb_dev_download": (C.c_int, [_P, _P, C.c_size_t]),
    "kb_dev_sync": (C.c_int, []),
    "kb_host_alloc_pinned": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "kb_host_free_pinned": (C.c_int, [_P]),
    "kb_timer_start": (C.c_int, []),
    "kb_timer_stop": (C.c_int, [_f32p]),
    "kb_flush_l2": (C.c_int, [C.c_size_t]),
    "kb_profiler_start": (C.c_int, []),
    "kb_profiler_stop": (C.c_int, []),
}
ABI_SYMBOLS = tuple(sorted(_SIGS))


def lib():
    """Load libkami_b200.so (fails loudly when it has not been built)."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise KamiError(-1, "libkami_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(p)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _ck(rc):
    if rc != 0:
        raise KamiError(rc, lib().kb_last_error().decode(errors="replace"))


def device_count():
    return lib().kb_device_count()


def init(device=-1):
    _ck(lib().kb_init(device))


def _fp(a):
    return a.ctypes.data_as(_f32p)


def _ip(a):
    return a.ctypes.data_as(_i32p)


def _vp(a):
    return a.ctypes.data_as(_P)


# ---- Env (kami/env.h) -------------------------------------------------------------------------
class Env:
    """Device-resident kami::Env.  Method names follow env.h."""

    def __init__(self):
        self.L = lib()
        self.h = _P()
        _ck(self.L.kb_env_create(C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kb_env_destroy(self.h)
            self.h = None

    def ply(self):
        v = C.c_int32()
        _ck(self.L.kb_env_ply(self.h, C.byref(v)))
        return v.value

    def turn(self):
        return 1.0 if self.ply() % 2 == 0 else -1.0

    def push(self, action):
        _ck(self.L.kb_env_push(self.h, int(action)))

    def pop(self):
        _ck(self.L.kb_env_pop(self.h))

    def actions(self):
        buf = np.zeros(MAX_ACTIONS, np.int32)
        n = C.c_int32()
        _ck(self.L.kb_env_actions(self.h, _ip(buf), MAX_ACTIONS, C.byref(n)))
        return buf[:n.value].copy()

    def observe(self):
        o = np.zeros(OBSIZE, np.float32)
        _ck(self.L.kb_env_observe(self.h, _fp(o)))
        return o

    def terminal(self):
        t, r, v = C.c_int32(), C.c_int32(), C.c_float()
        _ck(self.L.kb_env_terminal(self.h, C.byref(t), C.byref(v), C.byref(r)))
        return bool(t.value), v.value, r.value

    def encode(self, move):
        a = C.c_int32()
        _ck(self.L.kb_env_encode(self.h, int(move), C.byref(a)))
        return a.value

    def decode(self, action):
        m = C.c_int32()
        _ck(self.L.kb_env_decode(self.h, int(action), C.byref(m)))
        return m.value

    def bootstrap(self, window):
        v = C.c_float()
        _ck(self.L.kb_env_bootstrap(self.h, float(window), C.byref(v)))
        return v.value

    bootstrap_value = bootstrap

    def position(self):
        p = np.zeros(1, POSITION_DTYPE)
        _ck(self.L.kb_env_position(self.h, _vp(p)))
        return p


# ---- batched position kernels --------------------------------------------------------------------
def as_positions(raw):
    """[n,80] uint8 (oracle export) or structured array -> contiguous POSITION_DTYPE array."""
    a = np.ascontiguousarray(raw)
    if a.dtype != POSITION_DTYPE:
        a = a.reshape(-1, 80).view(POSITION_DTYPE).reshape(-1)
    return a


def encode_planes(positions):
    p = as_positions(positions)
    out = np.zeros((len(p), OBSIZE), np.float32)
    _ck(lib().kb_encode_planes(_vp(p), len(p), _fp(out)))
    return out


def encode_planes_tall(positions):
    """The tower's own input: bf16 planes in the swizzled tall-image layout (csrc/layout.cuh), as raw uint16 of shape
    [items][640 pixels][64 channels-in-swizzled-order]; device encoder kb_encode_planes_bf16_dev."""
    p = as_positions(positions)
    L = lib()
    nbytes = L.kb_net_planes_bytes(len(p))
    dpos, dpl = _P(), _P()
    _ck(L.kb_dev_alloc(C.byref(dpos), max(16, p.nbytes)))
    _ck(L.kb_dev_alloc(C.byref(dpl), nbytes))
    try:
        _ck(L.kb_dev_upload(dpos, _vp(p), p.nbytes))
        _ck(L.kb_encode_planes_bf16_dev(dpos, len(p), dpl))
        out = np.zeros(nbytes // 2, np.uint16)
        _ck(L.kb_dev_download(_vp(out), dpl, nbytes))
    finally:
        L.kb_dev_free(dpos)
        L.kb_dev_free(dpl)
    items = (len(p) + 6) // 7
    return out[: items * 640 * 64].reshape(items, 640, 64), out[items * 640 * 64:]


def legal_actions(positions):
    p = as_positions(positions)
    acts = np.zeros((len(p), MAX_ACTIONS), np.int32)
    cnt = np.zeros(len(p), np.int32)
    _ck(lib().kb_legal_actions(_vp(p), len(p), _ip(acts), _ip(cnt)))
    return acts, cnt


def apply_actions(positions, actions):
    p = as_positions(positions).copy()
    a = np.ascontiguousarray(actions, np.int32)
    _ck(lib().kb_apply_actions(_vp(p), len(p), _ip(a)))
    return p


def static_eval(positions):
    p = as_positions(positions)
    out = np.zeros(len(p), np.int32)
    _ck(lib().kb_static_eval(_vp(p), len(p), _ip(out)))
    return out


# ---- NN (kami/nn/nn.h) --------------------------------------------------------------------------
class Trainer:
    """One mini-batch of NN::train (kami/nn/nn.cpp:224-377) at a time: forward in training mode, loss,
    backward, plain SGD.  Weights travel as the flat fp32 blob of NN.load_blob (reference module names)."""

    def __init__(self, filters, residuals, max_batch):
        self.L = lib()
        self.filters, self.residuals = filters, residuals
        self.h = _P()
        _ck(self.L.kb_trainer_create(C.byref(self.h), filters, residuals, max_batch))
        self.n = self.L.kb_net_blob_floats(filters, residuals)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kb_trainer_destroy(self.h)
            self.h = None

    def load_blob(self, blob):
        blob = np.ascontiguousarray(blob, np.float32)
        _ck(self.L.kb_trainer_load_blob(self.h, _fp(blob), blob.size))

    def export_blob(self):
        out = np.zeros(self.n, np.float32)
        _ck(self.L.kb_trainer_export_blob(self.h, _fp(out), out.size))
        return out

    def export_grads(self):
        out = np.zeros(self.n, np.float32)
        _ck(self.L.kb_trainer_export_grads(self.h, _fp(out), out.size))
        return out

    def grad_buffer(self):
        """(device address, number of floats) of the flat gradient vector (for the NCCL all-reduce)."""
        p, n = C.c_void_p(), C.c_size_t()
        _ck(self.L.kb_trainer_grad_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def forward_backward(self, obs, obs_p, obs_v):
        obs = np.ascontiguousarray(obs, np.float32)
        obs_p = np.ascontiguousarray(obs_p, np.float32)
        obs_v = np.ascontiguousarray(obs_v, np.float32)
        loss = C.c_float()
        _ck(self.L.kb_trainer_forward_backward(self.h, _fp(obs), _fp(obs_p), _fp(obs_v), len(obs_v), C.byref(loss)))
        return loss.value

    def forward_backward_dev(self, obs_dev, obs_p_dev, obs_v_dev, batch, want_loss=True):
        loss = C.c_float()
        _ck(self.L.kb_trainer_forward_backward_dev(self.h, obs_dev, obs_p_dev, obs_v_dev, batch, C.byref(loss) if want_loss else None))
        return loss.value

    def apply_sgd(self, lr, grad_scale=1.0):
        _ck(self.L.kb_trainer_apply_sgd(self.h, float(lr), float(grad_scale)))

    def debug_activation(self, layer, which, board):
        out = np.zeros((256, 64), np.float32)
        ch = C.c_int()
        _ck(self.L.kb_trainer_debug_activation(self.h, layer, which, board, _fp(out), C.byref(ch)))
        return out[:ch.value].copy()


class DataParallelTrainer:
    """NN::train mini-batches data-parallel over the GPUs of the box from this one process (kb_dp_*): a Trainer replica
    per GPU, one NCCL all-reduce(sum) of the gradient bucket per step, the same SGD step everywhere."""

    def __init__(self, devices, filters, residuals, max_batch_per_device):
        self.L = lib()
        self.devices = list(devices)
        self.filters, self.residuals = filters, residuals
        dv = np.ascontiguousarray(self.devices, np.int32)
        self.h = _P()
        _ck(self.L.kb_dp_create(C.byref(self.h), _ip(dv), len(dv), filters, residuals, max_batch_per_device))
        self.n = self.L.kb_net_blob_floats(filters, residuals)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kb_dp_destroy(self.h)
            self.h = None

    def load_blob(self, blob):
        blob = np.ascontiguousarray(blob, np.float32)
        _ck(self.L.kb_dp_load_blob(self.h, _fp(blob), blob.size))

    def export_blob(self, rank=0):
        out = np.zeros(self.n, np.float32)
        _ck(self.L.kb_dp_export_blob(self.h, rank, _fp(out), out.size))
        return out

    def step(self, obs, obs_p, obs_v, lr, grad_scale=1.0):
        """obs [n_dev * b][1920], obs_p [n_dev * b][4672], obs_v [n_dev * b]; returns the summed loss."""
        obs = np.ascontiguousarray(obs, np.float32)
        obs_p = np.ascontiguousarray(obs_p, np.float32)
        obs_v = np.ascontiguousarray(obs_v, np.float32)
        loss = C.c_float()
        _ck(self.L.kb_dp_step(self.h, _fp(obs), _fp(obs_p), _fp(obs_v), len(obs_v) // len(self.devices), float(lr), float(grad_scale), C.byref(loss)))
        return loss.value


class NN:
    """kami::NN: NN(width, height, features, psize) with `filters` / `residuals` taken from the
    arguments instead of the global options map (nn.cpp:42-43)."""

    def __init__(self, filters=256, residuals=2, width=8, height=8, features=NFEATURES, psize=PSIZE):
        assert (width, height, features, psize) == (8, 8, NFEATURES, PSIZE)
        self.L = lib()
        self.filters, self.residuals = filters, residuals
        self.h = _P()
        self.generation = 0
        _ck(self.L.kb_net_create(C.byref(self.h), filters, residuals))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kb_net_destroy(self.h)
            self.h = None

    def isCUDA(self):
        return True

    def obsize(self):
        return OBSIZE

    def polsize(self):
        return PSIZE

    def get_generation(self):
        return self.generation

    def blob_floats(self):
        return self.L.kb_net_blob_floats(self.filters, self.residuals)

    def load_blob(self, blob):
        blob = np.ascontiguousarray(blob, np.float32)
        _ck(self.L.kb_net_load_blob(self.h, _fp(blob), blob.size))

    def infer(self, obs, batch=None):
        """NN::infer (nn.cpp:155-187): returns (policy [B,4672], value [B]) with the reference's
        value indexing."""
        obs = np.ascontiguousarray(obs, np.float32)
        B = batch or obs.size // OBSIZE
        pol = np.zeros((B, PSIZE), np.float32)
        val = np.zeros(B, np.float32)
        _ck(self.L.kb_net_infer(self.h, _fp(obs), B, _fp(pol), _fp(val)))
        return pol, val

    def forward_full(self, obs):
        obs = np.ascontiguousarray(obs, np.float32)
        B = obs.size // OBSIZE
        pol = np.zeros((B, PSIZE), np.float32)
        val = np.zeros((B, VALUE_WIDTH), np.float32)
        _ck(self.L.kb_net_forward_full(self.h, _fp(obs), B, _fp(pol), _fp(val)))
        return pol, val

    def debug_activation(self, which, board):
        out = np.zeros((256, 64), np.float32)
        ch = C.c_int32()
        _ck(self.L.kb_net_debug_activation(self.h, which, board, _fp(out), C.byref(ch)))
        return out[:ch.value].copy()

    def flops(self):
        t, h = C.c_double(), C.c_double()
        _ck(self.L.kb_net_flops(self.h, C.byref(t), C.byref(h)))
        return t.value, h.value


# ---- MCTS (kami/mcts.h) ----------------------------------------------------------------------------
def tree_cfg(**kw):
    c = TreeCfg()
    _ck(lib().kb_tree_default_cfg(C.byref(c)))
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


class TreePool:
    """n device-resident trees (the `MCTS trees[ibatch]` of selfplay.cpp:97)."""

    def __init__(self, n_trees, node_capacity=1 << 18, cfg=None):
        self.L = lib()
        self.cfg = cfg or tree_cfg()
        self.n = n_trees
        self.h = _P()
        _ck(self.L.kb_pool_create(C.byref(self.h), n_trees, node_capacity, C.byref(self.cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kb_pool_destroy(self.h)
            self.h = None

    def select(self):
        _ck(self.L.kb_pool_select(self.h))

    def leaf_positions(self):
        p = np.zeros(self.n, POSITION_DTYPE)
        _ck(self.L.kb_pool_leaf_positions(self.h, _vp(p)))
        return p

    def expand(self, policy, value, disable_bootstrap=False):
        policy = np.ascontiguousarray(policy, np.float32)
        value = np.ascontiguousarray(value, np.float32)
        assert policy.size == self.n * PSIZE and value.size == self.n
        _ck(self.L.kb_pool_expand(self.h, _fp(policy), _fp(value), int(disable_bootstrap)))

    def step(self, net, iters=1):
        _ck(self.L.kb_pool_step(self.h, net.h, iters))

    def step_hostio(self, net, iters, obs, policy, value):
        _ck(self.L.kb_pool_step_hostio(self.h, net.h, iters, _vp(obs), _vp(policy), _vp(value)))

    def step_hostio_compact(self, net, iters, leaves, prior, value):
        """step_hostio with 80-byte leaf positions and [128] legal-move priors on the wire."""
        _ck(self.L.kb_pool_step_hostio_compact(self.h, net.h, iters, _vp(leaves), _vp(prior), _vp(value)))

    def set_step_groups(self, groups):
        """step(): independent groups of trees, each with its own NN batch and stream (0 = default)."""
        _ck(self.L.kb_pool_set_step_groups(self.h, int(groups)))

    def set_profiling(self, on):
        """Phase timing of step() (phase_ms) is opt-in; off, step() records nothing and always fuses expand + select."""
        _ck(self.L.kb_pool_set_profiling(self.h, int(bool(on))))

    def set_terminal_cap(self, k):
        """A tree that absorbed k terminal visits in one step sits the step out (0 = off: the reference's batch)."""
        _ck(self.L.kb_pool_set_terminal_cap(self.h, int(k)))

    def set_split_select(self, mode):
        """step() with a terminal cap: planes before move lists, tower started under the move generators (-1 auto, 0 off, 1 on)."""
        _ck(self.L.kb_pool_set_split_select(self.h, int(mode)))

    def set_selfplay_nodes(self, nodes):
        _ck(self.L.kb_pool_set_selfplay_nodes(self.h, int(nodes)))

    def flush_trees(self):
        """flush_old_trees (selfplay.cpp:119-131): every tree back to the start position, partial trajectories dropped."""
        _ck(self.L.kb_pool_flush_trees(self.h))

    def leaf_actions(self):
        """Legal actions of every pending leaf in Env::actions() order: ([n,128] int32, -1 padded; [n] counts)."""
        a = np.zeros((self.n, MAX_ACTIONS), np.int32)
        c = np.zeros(self.n, np.int32)
        _ck(self.L.kb_pool_leaf_actions(self.h, _ip(a), _ip(c)))
        return a, c

    def expand_compact(self, prior, value, disable_bootstrap=False):
        prior = np.ascontiguousarray(prior, np.float32)
        value = np.ascontiguousarray(value, np.float32)
        assert prior.size == self.n * MAX_ACTIONS and value.size == self.n
        _ck(self.L.kb_pool_expand_compact(self.h, _fp(prior), _fp(value), int(disable_bootstrap)))

    def stats(self):
        s = PoolStats()
        _ck(self.L.kb_pool_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in PoolStats._fields_}

    def reset_stats(self):
        _ck(self.L.kb_pool_reset_stats(self.h))

    def set_policy_mode(self, dense):
        """0: softmax over the legal moves only (default); 1: dense softmax over all 4672 actions."""
        _ck(self.L.kb_pool_set_policy_mode(self.h, int(dense)))

    def request_game(self):
        """Selfplay::get_next_pgn: ask for the moves of the next game that finishes."""
        _ck(self.L.kb_pool_request_game(self.h))

    def take_game(self, cap=4096):
        """Action list of the requested finished game, or None while no game has finished."""
        buf = np.zeros(cap, np.int32)
        n = C.c_int()
        _ck(self.L.kb_pool_take_game(self.h, _vp(buf), cap, C.byref(n)))
        return buf[: n.value].copy() if n.value else None

    def set_hostio_groups(self, groups):
        """step_hostio pipelines: groups of trees with their own NN::infer batch and streams (0 = default)."""
        _ck(self.L.kb_pool_set_hostio_groups(self.h, int(groups)))

    def phase_ms(self):
        s = PhaseMs()
        _ck(self.L.kb_pool_last_phase_ms(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in PhaseMs._fields_}

    def drain_samples(self, max_samples):
        obs = np.zeros((max_samples, OBSIZE), np.float32)
        pi = np.zeros((max_samples, PSIZE), np.float32)
        z = np.zeros(max_samples, np.float32)
        n = C.c_int32()
        _ck(self.L.kb_pool_drain_samples(self.h, max_samples, _fp(obs), _fp(pi), _fp(z), C.byref(n)))
        return obs[:n.value], pi[:n.value], z[:n.value]

    def tree(self, i):
        return MCTS(pool=self, index=i)


class Arena:
    """Device-resident kami::eval (evaluate.cpp:10-160): `games` trees, two networks, one round per call."""

    def __init__(self, games, batch, nodes, colours, cfg=None):
        self.L = lib()
        self.games, self.batch, self.nodes = games, batch, nodes
        self.cfg = cfg or tree_cfg()
        col = np.ascontiguousarray(colours, np.int32)
        self.h = _P()
        _ck(self.L.kb_arena_create(C.byref(self.h), games, batch, nodes, C.byref(self.cfg), _ip(col), len(col)))
        self._ev = (ArenaGame * games)()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.kb_arena_destroy(self.h)
            self.h = None

    def _events(self, n):
        return [(self._ev[i].tree, self._ev[i].result, self._ev[i].colour) for i in range(n)]

    def round(self, current, candidate):
        """-> [(tree, result, colour)] of the games that ended, in the reference's order."""
        n = C.c_int32()
        _ck(self.L.kb_arena_round(self.h, current.h, candidate.h, self._ev, self.games, C.byref(n)))
        return self._events(n.value)

    def begin(self):
        """-> (finished games, current network's input rows, candidate's input rows)."""
        n, cn, dn = C.c_int32(), C.c_int32(), C.c_int32()
        cur = np.zeros((self.batch, OBSIZE), np.float32)
        cd = np.zeros((self.batch, OBSIZE), np.float32)
        _ck(self.L.kb_arena_begin(self.h, self._ev, self.games, C.byref(n), _fp(cur), C.byref(cn), _fp(cd), C.byref(dn)))
        return self._events(n.value), cur[:cn.value], cd[:dn.value]

    def end(self, cur_policy, cur_value, cd_policy, cd_value):
        arrs = [np.ascontiguousarray(a, np.float32) for a in (cur_policy, cur_value, cd_policy, cd_value)]
        _ck(self.L.kb_arena_end(self.h, *[_fp(a) if a.size else None for a in arrs]))

    def colours(self):
        out = np.zeros(self.games, np.int32)
        _ck(self.L.kb_arena_colours(self.h, _ip(out), self.games))
        return out

    def tree(self, i):
        """Read-only view of tree i (digest / n / root_children)."""
        view = TreePool.__new__(TreePool)
        view.L, view.h, view.n, view.cfg = self.L, None, self.games, self.cfg
        t = MCTS.__new__(MCTS)
        t.pool, t.i, t.L, t.ph = view, i, self.L, _P(self.L.kb_arena_pool(self.h))
        return t


class MCTS:
    """kami::MCTS call protocol (n / select / expand / pick / push / snapshot / reset) on one
    device-resident tree."""

    def __init__(self, cfg=None, pool=None, index=0, node_capacity=1 << 18):
        self.pool = pool or TreePool(1, node_capacity, cfg)
        self.i = index
        self.L = self.pool.L
        self.ph = self.pool.h

    def n(self):
        v = C.c_int32()
        _ck(self.L.kb_tree_n(self.ph, self.i, C.byref(v)))
        return v.value

    def select(self):
        """-> (need_eval, obs).  False means a terminal leaf was backed up (mcts.h:196-207)."""
        o = np.zeros(OBSIZE, np.float32)
        ne = C.c_int32()
        _ck(self.L.kb_tree_select(self.ph, self.i, _fp(o), C.byref(ne)))
        return bool(ne.value), o

    def expand(self, policy, value, disable_bootstrap=False):
        policy = np.ascontiguousarray(policy, np.float32)
        assert policy.size >= PSIZE
        _ck(self.L.kb_tree_expand(self.ph, self.i, _fp(policy), float(value), int(disable_bootstrap)))

    def pick(self, alpha=0.0, u01=0.0):
        a = C.c_int32()
        _ck(self.L.kb_tree_pick(self.ph, self.i, float(alpha), float(u01), C.byref(a)))
        return a.value

    def leaf_path(self):
        """Actions from the root to the leaf waiting for the network (empty when none is pending)."""
        a = np.zeros(255, np.int32)
        d = C.c_int32()
        _ck(self.L.kb_tree_leaf_path(self.ph, self.i, _ip(a), 255, C.byref(d)))
        return a[:d.value].copy()

    def push(self, action):
        _ck(self.L.kb_tree_push(self.ph, self.i, int(action)))

    def reset(self):
        _ck(self.L.kb_tree_reset(self.ph, self.i))

    def snapshot(self):
        o = np.zeros(PSIZE, np.float32)
        _ck(self.L.kb_tree_snapshot(self.ph, self.i, _fp(o)))
        return o

    def root_children(self):
        a = np.zeros(256, np.int32)
        n = np.zeros(256, np.int32)
        w = np.zeros(256, np.float32)
        p = np.zeros(256, np.float32)
        k = C.c_int32()
        _ck(self.L.kb_tree_root_children(self.ph, self.i, _ip(a), _ip(n), _fp(w), _fp(p), 256, C.byref(k)))
        k = k.value
        return a[:k].copy(), n[:k].copy(), w[:k].copy(), p[:k].copy()

    def root_w(self):
        v = C.c_float()
        _ck(self.L.kb_tree_root_w(self.ph, self.i, C.byref(v)))
        return v.value

    def digest(self):
        d, c = C.c_uint64(), C.c_int64()
        _ck(self.L.kb_tree_digest(self.ph, self.i, C.byref(d), C.byref(c)))
        return d.value, c.value

    def root_position(self):
        p = np.zeros(1, POSITION_DTYPE)
        _ck(self.L.kb_tree_env(self.ph, self.i, _vp(p)))
        return p
