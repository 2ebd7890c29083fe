"""Multi-GPU check (run under torchrun on >= 2 GPUs; not a pytest file):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_dist_check.py

Each rank calibrates the tiny ViT on ITS HALF of the golden's 8 images with the CUDA sweeps (NCCL all-reduce of the
FP64 per-candidate sums, exact distributed order statistics).  Checks: (1) every rank ends with bit-identical
parameters; (2) rank 0 then repeats the calibration alone on all images: top-1 agreement asserted, parameters
identical / total reported for the tiny goldens (4 images = 68 tokens per rank: not whole 32-token slabs) and
ASSERTED to be all for the BASELINE-sized cases (one DeiT-Tiny W4A4 / DeiT-Small W3A3 block, 32 images per rank),
where every FP32 partial is shard-invariant by construction."""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from adalog_b200.utils import dist as adist  # noqa: E402
from adalog_b200.utils import models as zoo  # noqa: E402
from adalog_b200.utils.calibrator import QuantCalibrator  # noqa: E402
from adalog_b200.utils.wrap_net import wrap_modules_in_net, wrap_reparamed_modules_in_net  # noqa: E402


def calibrate(g, images, dev):
    cfg = importlib.import_module(f'adalog_b200.configs.{g["bits"]}bit').Config()
    cfg.calib_size, cfg.calib_batch_size = images.shape[0], g['bs']
    model = zoo.create_model(g['model']).eval()
    model.load_state_dict(g['init_state'])
    model = wrap_modules_in_net(model.to(dev), cfg, reparam=True).to(dev).eval()
    loader = [(images[i:i + g['bs']], None) for i in range(0, images.shape[0], g['bs'])]
    cal = QuantCalibrator(model, loader)
    cal.progress = False
    cal.batching_quant_calib()
    return wrap_reparamed_modules_in_net(model)


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    cases = ['model_vit_test_w4a4', 'model_swin_test_w4a4', 'synthetic:deit_tiny_depth1_patch16_224:4',
             'synthetic:deit_small_depth1_patch16_224:3']
    for name in cases:
        if name.startswith('synthetic:'):
            # BASELINE-sized layers with 32 images per rank: every shard is a whole number of 32-token slabs, the
            # condition under which the FP32 partial sums -- and therefore every selected parameter -- must be
            # IDENTICAL to the single-process run (gemm_err.cu SLAB64, tests/test_gpu_invariance.py)
            _, mname, bits = name.split(':')
            torch.manual_seed(5)
            g = dict(model=mname, bits=int(bits), bs=32, init_state=zoo.create_model(mname).state_dict(),
                     images=torch.randn(32 * world, 3, 224, 224))
            must_match = True
        else:
            g = torch.load(os.path.join(ROOT, 'tests', 'golden', name + '.pt'), weights_only=False)
            must_match = False
        images = g['images'].to(dev)
        per = images.shape[0] // world
        model = calibrate(g, images[rank * per:(rank + 1) * per], dev)
        sd = model.state_dict()
        # (1) all ranks identical
        flat = torch.cat([v.detach().double().reshape(-1) for k, v in sorted(sd.items()) if 'quantizer' in k])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same_ranks = all(torch.equal(gathered[0], t) for t in gathered)
        dist.barrier()
        if rank == 0:
            real_active = adist.active
            adist.active = lambda: False            # single-process reference run on all images
            try:
                ref = calibrate(g, images, dev)
            finally:
                adist.active = real_active
            rsd = ref.state_dict()
            keys = [k for k in rsd if 'quantizer' in k]
            ident = sum(int(torch.equal(sd[k], rsd[k])) for k in keys)
            with torch.no_grad():
                probe = torch.cat([images, torch.randn(56, *images.shape[1:], device=dev,
                                                        generator=torch.Generator(dev).manual_seed(0))])
                a, b = model(probe), ref(probe)
            agree = (a.argmax(-1) == b.argmax(-1)).float().mean().item()
            print(f'[dist] {name}: world={world} ranks identical: {same_ranks}; vs single process: {ident}/{len(keys)} '
                  f'quantizer tensors bit-identical, logits rel diff {((a - b).norm() / b.norm()).item():.2e}, '
                  f'top-1 agreement {100 * agree:.1f}%')
            assert same_ranks
            if must_match:
                assert ident == len(keys) and agree == 1.0, \
                    'shards of whole 32-token slabs must reproduce the single-process result'
        dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
