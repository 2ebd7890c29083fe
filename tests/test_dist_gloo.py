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
