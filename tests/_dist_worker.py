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


def quantile_worker(rank, world, port, out_dir):
    """distributed order statistics (utils/dist.py kth_values, _fpcs.quantile_pair / chunked_quantile, post-GELU
    positive percentile) on this rank's shard of seeded data; rank 0 saves the results"""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from adalog_b200.quant_layers import _fpcs
    from adalog_b200.quant_layers.linear import PostGeluLogBasedBatchingQuantLinear as PG
    full = quantile_inputs()
    pct = torch.tensor([0.9, 1.0])
    out = {}
    x = full['cw']                                            # [rows, C] sharded along rows
    per = x.shape[0] // world
    up, lo = _fpcs.quantile_pair(x[rank * per:(rank + 1) * per], pct, 0)
    out['cw_up'], out['cw_lo'] = up, lo
    h = full['heads']                                         # [B, H, S, D] sharded along B, per-head statistics
    per = h.shape[0] // world
    hl = h[rank * per:(rank + 1) * per].transpose(0, 1).contiguous()
    out['head_up'], out['head_lo'] = _fpcs.chunked_quantile(hl.view(hl.shape[0], 1, -1), pct)
    t = full['tensor']                                        # per-tensor, with a small 2^24 stand-in (see the test)
    per = t.shape[0] // world
    out['t_up'], out['t_lo'] = _fpcs.chunked_quantile(t[rank * per:(rank + 1) * per].reshape(1, 1, -1), pct)
    g = full['gelu']
    per = g.shape[0] // world
    out['pos'] = PG._positive_percentile_dist(g[rank * per:(rank + 1) * per].reshape(-1), pct)
    if rank == 0:
        torch.save(out, os.path.join(out_dir, 'q.pt'))
    dist.barrier()
    dist.destroy_process_group()


def quantile_inputs():
    gen = torch.Generator().manual_seed(123)
    cw = torch.randn(8 * 50, 24, generator=gen) * torch.rand(24, generator=gen) * 3
    cw[::7, 3] = 0.0
    return {'cw': cw, 'heads': torch.randn(8, 3, 10, 6, generator=gen), 'tensor': torch.randn(8, 1000, generator=gen),
            'gelu': torch.nn.functional.gelu(torch.randn(8, 500, generator=gen))}


def quantile_worker_limit(rank, world, port, out_dir, limit):
    """quantile_worker with the 2^24 reduced-dimension limit of torch.quantile replaced by `limit` in chunked_quantile,
    so the three regimes of the chunk rule are reachable with small tensors"""
    from adalog_b200.quant_layers import _fpcs
    src = open(_fpcs.__file__).read().replace('(1 << 24)', f'({limit})')
    body = 'def chunked_quantile' + src.split('def chunked_quantile')[1]
    exec(compile(body, 'patched_fpcs', 'exec'), _fpcs.__dict__)
    quantile_worker(rank, world, port, out_dir)
