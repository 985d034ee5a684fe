"""TEST INFRASTRUCTURE ONLY -- the oracle is the checker, never the product.

PyTorch (CPU, fp32, autograd) restatement of ONE mini-batch of the reference's NN::train
(kami/nn/nn.cpp:224-377): NNModule::forward in training mode (BatchNorm2d with batch statistics,
momentum 0.1, eps 1e-5, running_var updated with the unbiased estimate -- LibTorch defaults,
nn.cpp:20-23, 45-56, 228), NNModule::loss (nn.cpp:93-105):

    loss = -sum(obs_p * log(p + 0.001))            (sum over the batch and the 4672 actions)
           + mean((v - obs_v)^2)                    (v is [B,256], obs_v is [B,1]: broadcast, mean over B*256)

backward, and one plain SGD step `w -= lr * grad` with lr = training_mlr / 1000 (nn.cpp:235-241, 353-354).
The arithmetic itself lives in LibTorch/ATen (unpinned third-party dependency, CMakeLists.txt:5;
torch 2.11.0 in this image).

PARITY PINNED: tests/test_train_oracle.py runs the unmodified reference NN::train (compiled on
LibTorch, oracle/_ref/libkami_ref_nn.so, one epoch, one mini-batch) on the same weights and batch
and compares every parameter and BatchNorm buffer after the step.
"""
import numpy as np
import torch
import torch.nn.functional as Fn

import nn_oracle as NO

BN_MOMENTUM = 0.1


def _id(x):
    return x


def _bf16(x):
    """Round a tensor to bf16 where the CUDA path stores it as bf16 (straight-through for autograd) and
    round the gradient that flows back into it as well (activation gradients are bf16 tensors there)."""
    y = x.detach().to(torch.bfloat16).to(torch.float32)
    out = x + (y - x).detach()
    if out.requires_grad:
        out.register_hook(lambda g: g.to(torch.bfloat16).to(torch.float32))
    return out


def _bf16_w(w):
    """bf16 operand copy of an fp32 master weight (gradient passes straight through to the master)."""
    return w + (w.detach().to(torch.bfloat16).to(torch.float32) - w).detach()


def _bn(x, p, name, new_stats):
    """BatchNorm2d in training mode on NCHW x; records the updated running statistics."""
    w, b = p[name + ".weight"], p[name + ".bias"]
    n = x.numel() // x.shape[1]
    mean = x.mean(dim=(0, 2, 3))
    var = x.var(dim=(0, 2, 3), unbiased=False)
    y = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + NO.BN_EPS)
    y = y * w[None, :, None, None] + b[None, :, None, None]
    with torch.no_grad():
        unbiased = var * (n / max(n - 1, 1))
        new_stats[name + ".running_mean"] = (1 - BN_MOMENTUM) * p[name + ".running_mean"] + BN_MOMENTUM * mean
        new_stats[name + ".running_var"] = (1 - BN_MOMENTUM) * p[name + ".running_var"] + BN_MOMENTUM * unbiased
    return y


def forward_train(p, obs, filters, residuals, new_stats, emulate_bf16=False, capture=None):
    """nn.cpp:59-91 with BatchNorm in training mode.  obs [B,64,30] (NHWC as Env::observe writes it).
    emulate_bf16: round conv weights, activations and activation gradients to bf16 at the points where
    the CUDA training step stores bf16 (a sharp check of the kernels; fp32 is the reference arithmetic)."""
    ra, rw = (_bf16, _bf16_w) if emulate_bf16 else (_id, _id)
    B = obs.shape[0]
    x = obs.reshape(B, 8, 8, NO.NFEATURES).permute(0, 3, 1, 2)

    def block(x, conv, bn, pad, skip=None):
        pre = ra(Fn.conv2d(x, rw(p[conv + ".weight"]), p[conv + ".bias"], padding=pad))
        y = torch.relu(_bn(pre, p, bn, new_stats))
        out = ra(y if skip is None else skip + y)
        if capture is not None:  # [B, C, 64] like kb_trainer_debug_activation
            capture.append((conv, pre.detach().flatten(2).numpy().copy(), out.detach().flatten(2).numpy().copy()))
        return out

    x = block(x, "conv1", "batchnorm1", 1)
    for i in range(residuals):
        r = "residual%d" % i
        skip = x
        x = block(x, r + ".conv1", r + ".batchnorm1", 1)
        x = block(x, r + ".conv2", r + ".batchnorm2", 1, skip)
    ph = block(x, "policyconv", "pbatchnorm", 0)
    ph = Fn.conv2d(ph, rw(p["policyconv2.weight"]), p["policyconv2.bias"])
    if emulate_bf16 and ph.requires_grad:
        ph.register_hook(lambda g: g.to(torch.bfloat16).to(torch.float32))  # dlogits are stored as bf16
    ph = ph.permute(0, 2, 3, 1).flatten(1)
    ph = torch.exp(torch.log_softmax(ph, 1))
    vh = torch.relu(_bn(Fn.conv2d(x, p["valueconv.weight"], p["valueconv.bias"]), p, "vbatchnorm", new_stats))
    vh = torch.tanh(Fn.linear(vh.flatten(1), p["valuefc.weight"], p["valuefc.bias"]))
    return ph, vh


def loss_fn(ph, vh, obs_p, obs_v):
    value_loss = Fn.mse_loss(vh, obs_v.reshape(-1, 1).expand_as(vh))  # mean over B*256 (broadcast, nn.cpp:96)
    policy_loss = -(obs_p * torch.log(ph + 0.001)).sum()
    return policy_loss + value_loss


def trainable(name):
    return not (name.endswith("running_mean") or name.endswith("running_var"))


def train_step(params, obs, obs_p, obs_v, filters, residuals, lr, emulate_bf16=False, capture=None):
    """One mini-batch of NN::train.  params: dict name -> numpy fp32 (nn_oracle.param_order).
    Returns (new params dict, loss, grads dict)."""
    p = {k: torch.tensor(np.asarray(v, np.float32), requires_grad=trainable(k)) for k, v in params.items()}
    new_stats = {}
    ph, vh = forward_train(p, torch.tensor(np.asarray(obs, np.float32)), filters, residuals, new_stats, emulate_bf16, capture)
    loss = loss_fn(ph, vh, torch.tensor(np.asarray(obs_p, np.float32)), torch.tensor(np.asarray(obs_v, np.float32)))
    loss.backward()
    out, grads = {}, {}
    for k, v in p.items():
        if trainable(k):
            g = v.grad if v.grad is not None else torch.zeros_like(v)
            grads[k] = g.numpy().copy()
            out[k] = (v.detach() - lr * g).numpy().copy()
        else:
            out[k] = new_stats[k].numpy().copy()
    return out, float(loss.item()), grads


def synthetic_targets(n, seed, legal_lists):
    """Replay-buffer shaped targets: a sparse visit distribution over each position's legal actions
    (MCTS::snapshot, mcts.h:341-348) and a result in {-1, draw value, +1} (selfplay.cpp:176-184)."""
    rng = np.random.RandomState(seed)
    pi = np.zeros((n, NO.PSIZE), np.float32)
    for i, acts in enumerate(legal_lists):
        w = rng.randint(0, 40, size=len(acts)).astype(np.float32)
        if w.sum() == 0:
            w[0] = 1
        pi[i, acts] = w / w.sum()
    z = rng.choice(np.array([-1.0, 0.5, 1.0], np.float32), size=n).astype(np.float32)
    return pi, z
