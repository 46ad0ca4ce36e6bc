"""PVNet weight container + drop-in `model.PVNet` (reference: 2_AlphaOmok/model.py:13-104).

PyTorch is used here only as the WEIGHT CONTAINER (state_dict / load_state_dict / parameters for the optimizer,
main.py:81-85,339-365) and for the out-of-scope training forward/backward.  Inference - i.e.
`model(x)` in eval mode under `torch.no_grad()` (agents.py:173-178, main.py:176-180, eval_main.py get_pv) - runs on
the hand-written sm_100a tower kernel (csrc/tower.cu) through the C ABI `ao_nn_forward`.

Parameter names are the reference's, so checkpoints (e.g. data/180927_9400_297233_step_model.pickle, which lacks
`num_batches_tracked`) load with the reference's own partial-update idiom.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _cabi


def seeded_state_dict(seed, n_block=10, inplanes=5, planes=128, board_size=9, bn_jitter=False, gain=1.0):
    """Reproducible "random-init PVNet" (BASELINE configs 1-4) without torch's RNG: numpy MT19937 draws with torch's
    default distributions (uniform +-1/sqrt(fan_in) for conv / linear weights and linear biases; BN gamma=1, beta=0,
    mean=0, var=1 as model.py:86-89), optionally with jittered BN statistics to exercise the BN folding.  Returns an
    ordered dict of float32 tensors with the reference's parameter names (load with strict=False: no
    num_batches_tracked keys, like the reference's 2018 checkpoints)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    a = board_size * board_size
    sd = {}

    def uni(shape, fan_in, g=1.0):
        b = g / np.sqrt(fan_in)
        return torch.from_numpy(rs.uniform(-b, b, size=shape).astype(np.float32))

    def bn(prefix, c):
        lo_hi = ((0.5, 1.5), (-0.2, 0.2), (-0.1, 0.1), (0.5, 1.5)) if bn_jitter else None
        for i, (name, const) in enumerate((("weight", 1.0), ("bias", 0.0), ("running_mean", 0.0), ("running_var", 1.0))):
            sd[prefix + "." + name] = (torch.from_numpy(rs.uniform(*lo_hi[i], c).astype(np.float32)) if bn_jitter
                                       else torch.full((c,), const))

    sd["conv1.weight"] = uni((planes, inplanes, 3, 3), inplanes * 9, gain)
    bn("bn1", planes)
    for i in range(n_block):
        for c in (1, 2):
            sd[f"layers.{i}.conv{c}.weight"] = uni((planes, planes, 3, 3), planes * 9, gain)
            bn(f"layers.{i}.bn{c}", planes)
    sd["policy_head.policy_head.weight"] = uni((2, planes, 1, 1), planes, gain)
    bn("policy_head.policy_bn", 2)
    sd["policy_head.policy_fc.weight"] = uni((a, 2 * a), 2 * a, gain)
    sd["policy_head.policy_fc.bias"] = uni((a,), 2 * a)
    sd["value_head.value_head.weight"] = uni((1, planes, 1, 1), planes, gain)
    bn("value_head.value_bn", 1)
    sd["value_head.value_fc1.weight"] = uni((planes, a), a, gain)
    sd["value_head.value_fc1.bias"] = uni((planes,), a)
    sd["value_head.value_fc2.weight"] = uni((1, planes), planes, gain)
    sd["value_head.value_fc2.bias"] = uni((1,), planes)
    return sd


def _conv(cin, cout, k):
    return nn.Conv2d(cin, cout, kernel_size=k, padding=k // 2, bias=False)


class ResBlock(nn.Module):
    """model.py:13-31: x -> relu(bn1(conv1 x)) -> bn2(conv2 .) + x -> relu"""

    def __init__(self, inplanes, planes):
        super().__init__()
        self.conv1, self.bn1 = _conv(inplanes, planes, 3), nn.BatchNorm2d(planes)
        self.conv2, self.bn2 = _conv(planes, planes, 3), nn.BatchNorm2d(planes)

    def forward(self, x):
        y = torch.relu(self.bn1(self.conv1(x)))
        return torch.relu(self.bn2(self.conv2(y)) + x)


class PolicyHead(nn.Module):
    """model.py:34-50: 1x1 conv -> 2 planes, BN, ReLU, channel-major flatten, FC 2A -> A, softmax over all cells"""

    def __init__(self, planes, board_size):
        super().__init__()
        a = board_size ** 2
        self.policy_head, self.policy_bn = _conv(planes, 2, 1), nn.BatchNorm2d(2)
        self.policy_fc = nn.Linear(2 * a, a)

    def forward(self, x):
        y = torch.relu(self.policy_bn(self.policy_head(x))).flatten(1)
        return torch.softmax(self.policy_fc(y), dim=-1)


class ValueHead(nn.Module):
    """model.py:53-73: 1x1 conv -> 1 plane, BN, ReLU, FC A -> planes, ReLU, FC planes -> 1, tanh"""

    def __init__(self, planes, board_size):
        super().__init__()
        a = board_size ** 2
        self.value_head, self.value_bn = _conv(planes, 1, 1), nn.BatchNorm2d(1)
        self.value_fc1, self.value_fc2 = nn.Linear(a, planes), nn.Linear(planes, 1)

    def forward(self, x):
        y = torch.relu(self.value_bn(self.value_head(x))).flatten(1)
        return torch.tanh(self.value_fc2(torch.relu(self.value_fc1(y)))).flatten()


class PVNet(nn.Module):
    """PVNet(n_block, inplanes, planes, board_size); forward(x[N,C,B,B]) -> (p[N,B*B], v[N])  (model.py:76-104)."""

    def __init__(self, n_block, inplanes, planes, board_size):
        super().__init__()
        self.n_block, self.inplanes, self.planes, self.board_size = n_block, inplanes, planes, board_size
        self.conv1, self.bn1 = _conv(inplanes, planes, 3), nn.BatchNorm2d(planes)
        self.layers = nn.Sequential(*[ResBlock(planes, planes) for _ in range(n_block)])
        self.policy_head = PolicyHead(planes, board_size)
        self.value_head = ValueHead(planes, board_size)
        for m in self.modules():  # model.py:86-89
            if isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
        self._ao_engine = None
        self._ao_fingerprint = None

    # ---- weights -> device engine (shared with ZeroAgent)
    def weights_fingerprint(self):
        """storage address + version counter per tensor (bumped by optimizer steps / load_state_dict); writes through
        `.data` bypass the counter - call `invalidate_inference_weights()` after those"""
        return tuple((t.data_ptr(), t._version) for t in self.state_dict().values())

    def invalidate_inference_weights(self):
        self._ao_fingerprint = None

    def _inference_engine(self, batch):
        fp = self.weights_fingerprint()
        if self._ao_engine is None or self._ao_engine.G < batch:
            if self._ao_engine is not None:
                self._ao_engine.close()
            self._ao_engine = _cabi.Engine(board_size=self.board_size, num_mcts=1, max_games=max(batch, 64),
                                           n_blocks=self.n_block, inplanes=self.inplanes, planes=self.planes,
                                           node_cap=4, device=_cabi.default_device(self))
            self._ao_fingerprint = None
        if self._ao_fingerprint != fp:
            self._ao_engine.load_state_dict(self.state_dict())
            self._ao_engine.choose_nn_precision()  # hi/lo split tower when these weights need it for 1e-4
            self._ao_fingerprint = fp
        return self._ao_engine

    def forward(self, x):
        if self.training or torch.is_grad_enabled():
            # training path (main.py:253-336): plain PyTorch, out of the hot-path scope
            h = self.layers(torch.relu(self.bn1(self.conv1(x))))
            return self.policy_head(h), self.value_head(h)
        eng = self._inference_engine(x.shape[0])
        p, v = eng.nn_forward(x.detach().float().cpu().numpy())
        return torch.from_numpy(p).to(x.device), torch.from_numpy(v).to(x.device)
