"""Host-side helpers with the reference's names and semantics (model/e2e_common.py).

Only what the three hot modules need: ``ModelBase.load_model`` (:22-38), ``to_cuda`` (:80-85),
``linear_tensor`` (:178-187), ``pad_list`` (:208-217), ``lecun_normal_init_parameters`` (:135-154).
"""
import math

import torch


class ModelBase(torch.nn.Module):
    """Mirror of model/e2e_common.py:15-67 (the parts the hot modules use)."""

    def forward(self, x):
        raise NotImplementedError

    @classmethod
    def load_model(cls, path, state_dict, opt=None):
        # same control flow as model/e2e_common.py:22-38 (keyword ``args=`` included)
        if path is not None:
            package = torch.load(path, map_location=lambda storage, loc: storage, weights_only=False)
            model = cls(args=package['opt'])
            if state_dict in package and package[state_dict] is not None:
                model.load_state_dict(package[state_dict])
        else:
            model = cls(opt)
        if opt is not None and len(opt.gpu_ids) > 0:
            model = model.cuda()
        return model

    @staticmethod
    def get_param_size(model):
        return sum(int(torch.tensor(p.size()).prod()) for p in model.parameters())


def to_cuda(m, x):
    """model/e2e_common.py:80-85: move x to the device of m's first parameter."""
    assert isinstance(m, torch.nn.Module)
    dev = next(m.parameters()).device
    if dev.type != 'cuda':
        return x
    return x.cuda(dev.index) if isinstance(x, torch.Tensor) else x


def linear_tensor(linear, x):
    """model/e2e_common.py:178-187."""
    y = linear(x.contiguous().view((-1, x.size()[-1])))
    return y.view((x.size()[:-1] + (-1,)))


def pad_list(xs, pad_value):
    """model/e2e_common.py:208-217."""
    n_batch = len(xs)
    max_len = max(x.size(0) for x in xs)
    pad = xs[0].new_zeros(n_batch, max_len, *xs[0].size()[1:]) + pad_value
    for i in range(n_batch):
        pad[i, :xs[i].size(0)] = xs[i]
    return pad


def lecun_normal_init_parameters(module):
    """model/e2e_common.py:135-154: N(0, 1/fan_in) weights, zero biases."""
    for p in module.parameters():
        data = p.data
        if data.dim() == 1:
            data.zero_()
        elif data.dim() == 2:
            n = data.size(1)
            data.normal_(0, 1. / math.sqrt(n))
        elif data.dim() == 4:
            n = data.size(1)
            for k in data.size()[2:]:
                n *= k
            data.normal_(0, 1. / math.sqrt(n))
        else:
            raise NotImplementedError
