"""N>1 path on CPU: two gloo ranks, each holding half of the calibration samples, must end with the parameters of the
single-process run (host logic: sharded capture tensors, all-reduced per-candidate scores, all-gathered order
statistics, identical top-k on every rank).  Scores come from the test-only oracle backend."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

import _dist_worker as W
from conftest import GOLDEN, load_golden

CASES = [('linear', 'linear_asym_w4a4'), ('cw', 'linear_cw_reparam_w3a3'), ('postgelu', 'linear_postgelu_w4a4'),
         ('qk', 'matmul_qk_a4'), ('pv', 'matmul_pv_s4a4')]


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('kind,name', CASES)
def test_two_ranks_match_single_process(kind, name, monkeypatch):
    g = load_golden(name)
    n = (g['x'] if 'x' in g else g['A']).shape[0]
    # single process with calib_batch_size = shard size: its per-batch partial sums are exactly the two shards' sums
    single = W.run_case(kind, g, 0, n, n // 2, monkeypatch)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(W.worker, args=(2, free_port(), kind, os.path.join(GOLDEN, name + '.pt'), d), nprocs=2, join=True)
        r0 = torch.load(os.path.join(d, 'rank0.pt'))
        r1 = torch.load(os.path.join(d, 'rank1.pt'))
    assert set(r0) == set(single)
    for k in single:
        assert torch.equal(r0[k], r1[k]), f'{k}: ranks disagree'
        assert torch.equal(r0[k], single[k]), f'{k}: differs from the single-process result'


@pytest.mark.parametrize('world,limit', [(2, 1 << 24), (4, 1 << 24), (2, 1000), (4, 4000), (4, 500)])
def test_distributed_order_statistics_are_exact(world, limit, monkeypatch):
    """Exact selection across ranks == torch.quantile / sort on the whole tensor, bit for bit: per-channel, per-head,
    per-tensor with the 2^24 chunk rule (shrunk to `limit` so that a chunk is a fraction of a rank's shard, one rank,
    or several ranks), and the post-GELU positive percentile."""
    from adalog_b200.quant_layers import _fpcs
    from adalog_b200.quant_layers.linear import PostGeluLogBasedBatchingQuantLinear as PG
    src = open(_fpcs.__file__).read()
    assert '(1 << 24)' in src
    full = W.quantile_inputs()
    pct = torch.tensor([0.9, 1.0])
    # single-process reference with the same (possibly shrunk) chunk limit
    mod = type(_fpcs)('fpcs_patched')
    mod.__dict__.update(_fpcs.__dict__)
    body = 'def chunked_quantile' + src.replace('(1 << 24)', f'({limit})').split('def chunked_quantile')[1]
    exec(compile(body, 'patched_fpcs', 'exec'), mod.__dict__)
    ref = {}
    ref['cw_up'], ref['cw_lo'] = mod.quantile_pair(full['cw'], pct, 0)
    h = full['heads'].transpose(0, 1).contiguous()
    ref['head_up'], ref['head_lo'] = mod.chunked_quantile(h.view(h.shape[0], 1, -1), pct)
    ref['t_up'], ref['t_lo'] = mod.chunked_quantile(full['tensor'].reshape(1, 1, -1), pct)
    ref['pos'] = PG.positive_percentile(full['gelu'].reshape(-1), pct)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(W.quantile_worker_limit, args=(world, free_port(), d, limit), nprocs=world, join=True)
        got = torch.load(os.path.join(d, 'q.pt'))
    for k in ref:
        assert torch.equal(got[k], ref[k]), (k, got[k], ref[k])
