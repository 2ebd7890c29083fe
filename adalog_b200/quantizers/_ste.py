"""Straight-through estimators (reference quantizers/_ste.py:5-14); only live in training_mode (BRECQ)."""
import torch


def round_ste(x: torch.Tensor):
    return (x.round() - x).detach() + x


def floor_ste(x: torch.Tensor):
    return (x.floor() - x).detach() + x


def ceil_ste(x: torch.Tensor):
    return (x.ceil() - x).detach() + x


def flag(t):
    """bool(t) for a 0-d flag buffer (e.g. `bias_reparamed`) without a device synchronisation on every inference
    forward: under no_grad the value is read once and remembered on the tensor until it is modified in place or moved
    (writes through `.data` do not bump the version counter: reparam_bias() invalidates explicitly).  With autograd
    enabled (BRECQ training drives these modules) the tensor is always read."""
    if torch.is_grad_enabled():
        return bool(t)
    key = (t.data_ptr(), t._version)
    hit = getattr(t, '_adalog_flag', None)
    if hit is not None and hit[0] == key:
        return hit[1]
    v = bool(t)
    try:
        t._adalog_flag = (key, v)
    except AttributeError:
        pass
    return v
