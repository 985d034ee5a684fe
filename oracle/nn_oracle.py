"""TEST INFRASTRUCTURE ONLY -- the oracle is the checker, never the product.

numpy fp32 restatement of the reference network forward (kami/nn/nn.cpp:26-34 residual
block, :59-91 NNModule::forward, :155-187 NN::infer).  The arithmetic itself lives in
LibTorch/ATen (unpinned third-party dependency, CMakeLists.txt:5; torch 2.11.0 in this
image): Conv2d 3x3 zero-pad / 1x1 with bias, BatchNorm2d in eval mode (eps 1e-5),
ReLU, log_softmax+exp over all 4672 logits, Linear(64,256), tanh.

PARITY PINNED: tests/test_nn_oracle.py checks this file against the unmodified reference
NN compiled on LibTorch (oracle/_ref/libkami_ref_nn.so) and against the committed
fixtures tests/golden/nn_*.npz that tests/golden/make_golden.py generated from it.
"""
import numpy as np

NFEATURES = 30
PSIZE = 4672
BN_EPS = 1e-5


def param_order(filters, residuals):
    """(name, shape) in the order of the flat fp32 weight blob that
    kb_net_load_blob (include/kami_b200.h) consumes.  Names are the reference's
    register_module names (nn.cpp:20-23, 45-56)."""
    F = filters
    out = []

    def conv(name, co, ci, k):
        out.append((name + ".weight", (co, ci, k, k)))
        out.append((name + ".bias", (co,)))

    def bn(name, c):
        for s in ("weight", "bias", "running_mean", "running_var"):
            out.append((name + "." + s, (c,)))

    conv("conv1", F, NFEATURES, 3)
    bn("batchnorm1", F)
    for i in range(residuals):
        r = "residual%d" % i
        conv(r + ".conv1", F, F, 3)
        bn(r + ".batchnorm1", F)
        conv(r + ".conv2", F, F, 3)
        bn(r + ".batchnorm2", F)
    conv("policyconv", 128, F, 1)
    bn("pbatchnorm", 128)
    conv("policyconv2", 73, 128, 1)
    conv("valueconv", 1, F, 1)
    bn("vbatchnorm", 1)
    out.append(("valuefc.weight", (256, 64)))
    out.append(("valuefc.bias", (256,)))
    return out


def init_params(filters, residuals, seed, randomize_bn=True):
    """Random-init weights of the reference architecture.  Same distributions as LibTorch's
    defaults (kaiming-uniform a=sqrt(5) == U(+-1/sqrt(fan_in)) for weights and biases); with
    randomize_bn the BatchNorm affine/running stats are also randomised so that BN folding is
    actually exercised (LibTorch's init is the identity: gamma 1, beta 0, mean 0, var 1)."""
    rng = np.random.RandomState(seed)
    p = {}
    for name, shape in param_order(filters, residuals):
        leaf = name.rsplit(".", 1)[1]
        mod = name.rsplit(".", 1)[0]
        if "batchnorm" in mod:
            if leaf == "weight":
                v = rng.uniform(0.5, 1.5, shape) if randomize_bn else np.ones(shape)
            elif leaf == "bias":
                v = rng.uniform(-0.3, 0.3, shape) if randomize_bn else np.zeros(shape)
            elif leaf == "running_mean":
                v = rng.uniform(-0.2, 0.2, shape) if randomize_bn else np.zeros(shape)
            else:
                v = rng.uniform(0.5, 2.0, shape) if randomize_bn else np.ones(shape)
        else:
            if leaf == "weight":
                fan_in = int(np.prod(shape[1:]))
                p["_fan_" + mod] = fan_in
            fan_in = p["_fan_" + mod]
            b = 1.0 / np.sqrt(fan_in)
            v = rng.uniform(-b, b, shape)
        p[name] = v.astype(np.float32)
    return {k: v for k, v in p.items() if not k.startswith("_")}


def pack_blob(params, filters, residuals):
    return np.concatenate([np.asarray(params[n], np.float32).reshape(-1) for n, _ in param_order(filters, residuals)])


def unpack_blob(blob, filters, residuals):
    p, o = {}, 0
    for n, s in param_order(filters, residuals):
        k = int(np.prod(s))
        p[n] = np.asarray(blob[o:o + k], np.float32).reshape(s)
        o += k
    assert o == len(blob)
    return p


def _conv(x, w, b):
    """x [B,C,8,8], w [O,C,k,k] (k 1 or 3, zero pad k//2), fp32."""
    B, Cc, H, W = x.shape
    O, _, k, _ = w.shape
    if k == 1:
        y = np.einsum("bchw,oc->bohw", x, w[:, :, 0, 0], optimize=True)
    else:
        xp = np.zeros((B, Cc, H + 2, W + 2), np.float32)
        xp[:, :, 1:-1, 1:-1] = x
        cols = np.stack([xp[:, :, dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)], axis=2)
        y = np.einsum("bckhw,ock->bohw", cols, w.reshape(O, Cc, 9), optimize=True)
    return (y + b[None, :, None, None]).astype(np.float32)


def _bn(x, p, name):
    g, b = p[name + ".weight"], p[name + ".bias"]
    m, v = p[name + ".running_mean"], p[name + ".running_var"]
    s = (g / np.sqrt(v + np.float32(BN_EPS))).astype(np.float32)
    return ((x - m[None, :, None, None]) * s[None, :, None, None] + b[None, :, None, None]).astype(np.float32)


def forward(params, obs, residuals=None):
    """obs [B,8,8,30] fp32 (Env::observe layout) -> (policy [B,4672], value [B,256])."""
    p = params
    if residuals is None:
        residuals = len([k for k in p if k.endswith(".conv1.weight") and k.startswith("residual")])
    x = np.ascontiguousarray(np.asarray(obs, np.float32).reshape(-1, 8, 8, NFEATURES).transpose(0, 3, 1, 2))
    x = np.maximum(_bn(_conv(x, p["conv1.weight"], p["conv1.bias"]), p, "batchnorm1"), 0)
    for i in range(residuals):
        r = "residual%d" % i
        y = np.maximum(_bn(_conv(x, p[r + ".conv1.weight"], p[r + ".conv1.bias"]), p, r + ".batchnorm1"), 0)
        y = np.maximum(_bn(_conv(y, p[r + ".conv2.weight"], p[r + ".conv2.bias"]), p, r + ".batchnorm2"), 0)
        x = x + y  # ReLU before the add, none after (nn.cpp:31)
    ph = np.maximum(_bn(_conv(x, p["policyconv.weight"], p["policyconv.bias"]), p, "pbatchnorm"), 0)
    ph = _conv(ph, p["policyconv2.weight"], p["policyconv2.bias"])  # [B,73,8,8]
    ph = ph.transpose(0, 2, 3, 1).reshape(len(x), PSIZE)  # index = 73*sq + atype (nn.cpp:78-79)
    ph = ph - ph.max(axis=1, keepdims=True)
    e = np.exp(ph.astype(np.float64))
    policy = (e / e.sum(axis=1, keepdims=True)).astype(np.float32)
    vh = np.maximum(_bn(_conv(x, p["valueconv.weight"], p["valueconv.bias"]), p, "vbatchnorm"), 0)
    vh = vh.reshape(len(x), 64)
    value = np.tanh(vh @ p["valuefc.weight"].T + p["valuefc.bias"]).astype(np.float32)
    return policy, value


def infer(params, obs):
    """NN::infer semantics (nn.cpp:155-187) including the value memcpy quirk:
    value[i] = vh.flat[i], i.e. vh[i // 256][i % 256]."""
    policy, v256 = forward(params, obs)
    B = len(policy)
    return policy, v256.reshape(-1)[:B].copy()


def flops_per_position(filters, residuals):
    """Dense-convention FLOPs (2*64*K*N per conv), SURVEY.md section 8(d)."""
    F = filters
    tower = 2 * 64 * (9 * NFEATURES) * F + residuals * 2 * (2 * 64 * 9 * F * F)
    heads = 2 * 64 * F * 128 + 2 * 64 * 128 * 73 + 2 * 64 * F * 1 + 2 * 64 * 256
    return tower, heads
