"""CPU ORACLE (test infrastructure, NOT product code): fp32 PyTorch restatement of PVNet.forward in eval mode.

Follows /root/reference/2_AlphaOmok/model.py:13-104 (ResBlock :22-31, PolicyHead :43-50, ValueHead :63-73,
PVNet.forward :97-104) as a pure function of a state_dict, so it runs without the reference on the GPU box.
Pinned against the real `model.PVNet` by tests/golden/make_golden.py (nn_*.npz fixtures).

Also holds the deterministic weight generator used by tests and bench ("random-init PVNet" with a numpy seed, so
the weights are reproducible without torch's RNG): same distributions as torch's default init
(uniform +-1/sqrt(fan_in) for conv / linear weights and linear biases; BN gamma=1, beta=0, mean=0, var=1 as
model.py:86-89), optionally with jittered BN statistics to exercise the BN folding.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def state_dict_keys(n_block: int):
    keys = ["conv1.weight"] + [f"bn1.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
    for i in range(n_block):
        for c in (1, 2):
            keys.append(f"layers.{i}.conv{c}.weight")
            keys += [f"layers.{i}.bn{c}.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
    keys.append("policy_head.policy_head.weight")
    keys += [f"policy_head.policy_bn.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
    keys += ["policy_head.policy_fc.weight", "policy_head.policy_fc.bias", "value_head.value_head.weight"]
    keys += [f"value_head.value_bn.{s}" for s in ("weight", "bias", "running_mean", "running_var")]
    keys += ["value_head.value_fc1.weight", "value_head.value_fc1.bias",
             "value_head.value_fc2.weight", "value_head.value_fc2.bias"]
    return keys


def make_state_dict(seed: int, n_block=10, inplanes=5, planes=128, board_size=9, bn_jitter=False, gain=1.0):
    """Deterministic (numpy MT19937) random-init PVNet weights as an ordered dict of float32 torch tensors."""
    rs = np.random.RandomState(seed)
    A = board_size * board_size
    sd = {}

    def uni(shape, fan_in, g=1.0):
        b = g / np.sqrt(fan_in)
        return torch.from_numpy(rs.uniform(-b, b, size=shape).astype(np.float32))

    def bn(prefix, c):
        if bn_jitter:
            sd[prefix + ".weight"] = torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32))
            sd[prefix + ".bias"] = torch.from_numpy(rs.uniform(-0.2, 0.2, c).astype(np.float32))
            sd[prefix + ".running_mean"] = torch.from_numpy(rs.uniform(-0.1, 0.1, c).astype(np.float32))
            sd[prefix + ".running_var"] = torch.from_numpy(rs.uniform(0.5, 1.5, c).astype(np.float32))
        else:
            sd[prefix + ".weight"] = torch.ones(c)
            sd[prefix + ".bias"] = torch.zeros(c)
            sd[prefix + ".running_mean"] = torch.zeros(c)
            sd[prefix + ".running_var"] = torch.ones(c)

    sd["conv1.weight"] = uni((planes, inplanes, 3, 3), inplanes * 9, gain)
    bn("bn1", planes)
    for i in range(n_block):
        for c in (1, 2):
            sd[f"layers.{i}.conv{c}.weight"] = uni((planes, planes, 3, 3), planes * 9, gain)
            bn(f"layers.{i}.bn{c}", planes)
    sd["policy_head.policy_head.weight"] = uni((2, planes, 1, 1), planes, gain)
    bn("policy_head.policy_bn", 2)
    sd["policy_head.policy_fc.weight"] = uni((A, 2 * A), 2 * A, gain)
    sd["policy_head.policy_fc.bias"] = uni((A,), 2 * A)
    sd["value_head.value_head.weight"] = uni((1, planes, 1, 1), planes, gain)
    bn("value_head.value_bn", 1)
    sd["value_head.value_fc1.weight"] = uni((planes, A), A, gain)
    sd["value_head.value_fc1.bias"] = uni((planes,), A)
    sd["value_head.value_fc2.weight"] = uni((1, planes), planes, gain)
    sd["value_head.value_fc2.bias"] = uni((1,), planes)
    return sd


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], training=False, eps=BN_EPS)


def n_blocks_of(sd) -> int:
    n = 0
    while f"layers.{n}.conv1.weight" in sd:
        n += 1
    return n


@torch.no_grad()
def pvnet_forward(sd, x: torch.Tensor):
    """x float32 [N,C,B,B] -> (p float32 [N,A] softmax over all cells, v float32 [N] tanh)."""
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    h = F.relu(_bn(F.conv2d(x, sd["conv1.weight"], padding=1), sd, "bn1"))
    for i in range(n_blocks_of(sd)):
        r = h
        o = F.relu(_bn(F.conv2d(h, sd[f"layers.{i}.conv1.weight"], padding=1), sd, f"layers.{i}.bn1"))
        o = _bn(F.conv2d(o, sd[f"layers.{i}.conv2.weight"], padding=1), sd, f"layers.{i}.bn2")
        h = F.relu(o + r)
    p = F.relu(_bn(F.conv2d(h, sd["policy_head.policy_head.weight"]), sd, "policy_head.policy_bn"))
    p = p.reshape(p.shape[0], -1)
    p = F.softmax(F.linear(p, sd["policy_head.policy_fc.weight"], sd["policy_head.policy_fc.bias"]), dim=-1)
    v = F.relu(_bn(F.conv2d(h, sd["value_head.value_head.weight"]), sd, "value_head.value_bn"))
    v = v.reshape(v.shape[0], -1)
    v = F.relu(F.linear(v, sd["value_head.value_fc1.weight"], sd["value_head.value_fc1.bias"]))
    v = torch.tanh(F.linear(v, sd["value_head.value_fc2.weight"], sd["value_head.value_fc2.bias"]))
    return p, v.reshape(-1)
