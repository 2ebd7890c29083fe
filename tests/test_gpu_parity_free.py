"""-m gpu: free-running calibration parity against the reference-on-GPU at BASELINE.json's sizes (tests/gpu_parity.py):
config 1 in full (DeiT-Tiny W4A4, 32 images, 3150 evaluations) and the patch-embedding + block 0 + head slice of
config 2 (DeiT-Small W3A3, 128 images).  Asserts: per-candidate scores within 1e-5 of the reference on identical
candidates (3e-5 for the split-3 patch embedding), forced checkpoint and logits bit-identical, every top-k set that
differs lies inside the reference's own FP32-vs-FP64 noise, exact ties stay exact ties, top-1 agreement 100 %."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(rec):
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f'parity_{rec["model"]}_w{rec["bits"]}_{rec["images"]}img.json'), 'w') as f:
        json.dump(rec, f, indent=1)
    assert rec['forced_max_rel_diff'] <= 1e-5, rec
    assert rec['forced_max_rel_diff_patch_embed'] <= 3e-5, rec
    assert rec['forced_checkpoint_bit_identical'] and rec['forced_logits_bit_identical'], rec
    assert rec['exact_ties_preserved'] == rec['exact_ties_in_reference'], rec
    assert rec['topk_sets_differing_outside_reference_noise'] == 0, rec
    assert rec['free_running_top1_agreement'] == 1.0, rec


def test_config1_deit_tiny_w4a4_32img_full():
    from gpu_parity import run_parity
    _check(run_parity('deit_tiny_patch16_224', 4, 32))


def test_config2_slice_deit_small_w3a3_128img():
    from gpu_parity import run_parity
    _check(run_parity('deit_small_depth1_patch16_224', 3, 128))
