"""Straight-through estimators (reference quantizers/_ste.py:5-14); only live in training_mode (BRECQ)."""
import torch


def round_ste(x: torch.Tensor):
    return (x.round() - x).detach() + x


def floor_ste(x: torch.Tensor):
    return (x.floor() - x).detach() + x


def ceil_ste(x: torch.Tensor):
    return (x.ceil() - x).detach() + x


def assign(p, value):
    """p <- value in place THROUGH the tensor itself (not `.data`), so that p._version is bumped: the caches keyed on
    (data_ptr, _version) -- round_ste(zero_point) in ops.py, the tensor-core weight operand in sweep.py, flag() below --
    see every write this package makes to quantizer state, weights and biases.  (External code that writes through
    `.data`, as the reference's BRECQ does, must call `invalidate_caches(module)`.)"""
    with torch.no_grad():
        p.copy_(value.to(p.device) if isinstance(value, torch.Tensor) else value)
    return p


def invalidate_caches(module):
    """Forget every derived tensor cached on `module` and its quantizers (after an external write through `.data`)."""
    for m in module.modules():
        m.__dict__.pop('_tc_cache', None)
        m.__dict__.pop('_pct_cache', None)
        ctx = m.__dict__.get('_ctx')               # fixed operands of a search in progress (sweep._cached_fixed)
        if ctx is not None:
            ctx.__dict__.pop('_fixed_cache', None)
        for t in list(m.parameters(recurse=False)) + list(m.buffers(recurse=False)):
            for attr in ('_adalog_zr', '_adalog_flag'):
                if hasattr(t, attr):
                    try:
                        delattr(t, attr)
                    except AttributeError:
                        pass


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
