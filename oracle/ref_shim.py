"""Loader for the UNMODIFIED reference (GoatWu/AdaLog at /root/reference) on a CPU-only box.

TEST INFRASTRUCTURE ONLY.  Used exclusively by oracle/make_golden.py (run in the build
container, where /root/reference exists) to generate the fixtures under tests/golden/.
Nothing in the product (adalog_b200/), bench.py or the -m gpu tests imports this file.

The reference refuses to run without CUDA (quant_layers/linear.py:113-117, matmul.py:97-101,
conv.py:145-149) and calls .cuda() everywhere; the shim below makes those identity ops so the
reference's own arithmetic runs on torch-CPU.  No reference file is modified or copied.
"""
import sys
import torch

REF_ROOT = '/root/reference'


class _Props:
    total_memory = 16 * 2 ** 30  # only drives parallel_eq_n (candidate chunking)


def install(total_memory=None):
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if total_memory is not None:
        _Props.total_memory = total_memory
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    torch.cuda.get_device_properties = lambda *a, **k: _Props()
