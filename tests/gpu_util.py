"""helpers shared by the -m gpu parity tests"""
import torch

RTOL = 1e-5   # north_star: per-candidate errors agree within 1e-5 relative in FP32


def rel_diff(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs() / b.abs().clamp_min(1e-30)).max().item()


def assert_sims_close(got, ref, what='', rtol=RTOL):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    rd = rel_diff(got, ref)
    assert rd <= rtol, f'{what}: max relative difference {rd:.3e} > {rtol:.1e}'
    return rd


def near_tie_ok(sims_ref, idx_ref, idx_got, dim, k, rtol=4 * RTOL):
    """A differing top-k is acceptable only if every extra pick is within rtol of the reference's k-th best score."""
    if torch.equal(idx_ref, idx_got):
        return True
    s = sims_ref.double()
    d = dim % s.dim()
    vals_ref = torch.gather(s, d, idx_ref.reshape([k if i == d else n for i, n in enumerate(s.shape)]))
    vals_got = torch.gather(s, d, idx_got.reshape([k if i == d else n for i, n in enumerate(s.shape)]))
    kth = vals_ref.min(dim=d, keepdim=True).values
    slack = kth.abs() * rtol
    return bool((vals_got >= kth - slack).all())


class ForcedTopk:
    """Record the product's similarity tensors while replaying the oracle's selections (teacher forcing), so that
    every evaluation of a search is compared on identical candidates."""

    def __init__(self, oracle_evals):
        self.oracle = oracle_evals
        self.got = []
        self._orig = torch.topk
        self._orig_argmax = torch.Tensor.argmax

    def __enter__(self):
        def forced(inp, k, dim=-1, **kw):
            i = len(self.got)
            o = self.oracle[i]
            assert o['k'] == k
            own = self._orig(inp, k=k, dim=dim, **kw)[1]
            self.got.append(dict(sims=inp.detach().clone(), idx=own, k=k, dim=dim))
            idx = o['idx'].reshape(own.shape)
            return torch.gather(inp, dim, idx), idx
        torch.topk = forced
        if any(e.get('argmax') for e in self.oracle):
            # the twin-uniform search selects with Tensor.argmax (linear.py:691): same forcing
            orig = self._orig_argmax

            def forced_argmax(t, *a, **kw):
                i = len(self.got)
                if i >= len(self.oracle) or not self.oracle[i].get('argmax'):
                    return orig(t, *a, **kw)
                own = orig(t, *a, **kw)
                self.got.append(dict(sims=t.detach().clone(), idx=own, k=1, dim=0))
                return self.oracle[i]['idx'].reshape(own.shape)
            torch.Tensor.argmax = forced_argmax
        return self

    def __exit__(self, *a):
        torch.topk = self._orig
        torch.Tensor.argmax = self._orig_argmax

    def report(self, what, rtol=None, rtol_by_eval=None):
        worst, flips, bad = 0.0, 0, 0
        assert len(self.got) == len(self.oracle), (len(self.got), len(self.oracle))
        for i, (g, o) in enumerate(zip(self.got, self.oracle)):
            worst = max(worst, assert_sims_close(g['sims'], o['sims'].reshape(g['sims'].shape), f'{what} eval {i}',
                                                 (rtol_by_eval or {}).get(i, rtol or RTOL)))
            if not torch.equal(g['idx'], o['idx'].reshape(g['idx'].shape)):
                flips += 1
                if not near_tie_ok(o['sims'].reshape(g['sims'].shape), o['idx'].reshape(g['idx'].shape), g['idx'],
                                   g['dim'], g['k']):
                    bad += 1
        print(f'[parity] {what}: {len(self.got)} evals, max rel diff {worst:.2e}, '
              f'{flips} top-k lists differ (all near-ties: {bad == 0})')
        assert bad == 0, f'{what}: {bad} selections differ beyond the near-tie margin'
        return worst, flips
