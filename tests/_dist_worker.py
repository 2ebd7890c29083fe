"""worker for tests/test_dist_gloo.py (spawned, world_size 2, gloo, CPU)"""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)


class _Patch:
    """minimal monkeypatch stand-in for spawned processes"""

    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def run_case(kind, g, lo, hi, bs, patch=None):
    import _oracle_backend as fake
    from adalog_b200 import quant_layers as QL
    fake.install(patch or _Patch(), bs, g['cfg'].get('memory', 8 * 2 ** 30))
    c = g['cfg']
    with torch.no_grad():
        if kind == 'linear' or kind == 'postgelu' or kind == 'cw':
            cls = {'linear': QL.AsymmetricallyBatchingQuantLinear, 'postgelu': QL.PostGeluLogBasedBatchingQuantLinear,
                   'cw': QL.AsymmetricallyChannelWiseBatchingQuantLinear}[kind]
            extra = dict(quantizer='adalog') if kind == 'postgelu' else {}
            m = cls(c['in_f'], c['out_f'], bias=c['bias'], w_bit=c['w_bit'], a_bit=c['a_bit'], calib_batch_size=bs,
                    eq_n=128, fpcs=True, steps=6, search_round=3, n_V=c['n_V'], **extra)
            m.weight.data.copy_(g['weight'])
            if c['bias']:
                m.bias.data.copy_(g['bias'])
            ln = None
            if kind == 'cw':
                ln = torch.nn.LayerNorm(c['in_f'])
                ln.weight.data.copy_(g['ln_weight'])
                ln.bias.data.copy_(g['ln_bias'])
                m.prev_layer = ln
            m.raw_input, m.raw_out = g['x'][lo:hi].clone(), g['raw_out'][lo:hi].clone()
            m.hyperparameter_searching()
            if ln is not None:
                m.reparam()
        else:
            kw = dict(A_bit=c['A_bit'], B_bit=c['B_bit'], calib_batch_size=bs, search_round=3, eq_n=128,
                      head_channel_wise=c['hcw'], num_heads=c['H'], fpcs=True, steps=6)
            m = (QL.PostSoftmaxAsymmetricallyBatchingQuantMatMul(quantizer='adalog', **kw) if kind == 'pv'
                 else QL.AsymmetricallyBatchingQuantMatMul(**kw))
            m.raw_input, m.raw_out = [g['A'][lo:hi].clone(), g['B'][lo:hi].clone()], g['raw_out'][lo:hi].clone()
            m.hyperparameter_searching()
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def worker(rank, world, port, kind, golden_path, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    g = torch.load(golden_path, weights_only=False)
    n = (g['x'] if 'x' in g else g['A']).shape[0]
    per = n // world
    sd = run_case(kind, g, rank * per, (rank + 1) * per, per)
    torch.save(sd, os.path.join(out_dir, f'rank{rank}.pt'))
    dist.barrier()
    dist.destroy_process_group()
