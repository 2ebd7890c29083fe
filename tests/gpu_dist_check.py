"""Multi-GPU check (run under torchrun on >= 2 GPUs; not a pytest file):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/gpu_dist_check.py

Each rank calibrates the tiny ViT on ITS HALF of the golden's 8 images with the CUDA sweeps (NCCL all-reduce of the
FP64 per-candidate sums, all-gathered order statistics).  Checks: (1) every rank ends with bit-identical parameters;
(2) rank 0 then repeats the calibration alone on all 8 images: the sharded result must match it (FP64 sums are
reassociated across ranks, so equality is expected up to exact-tie-free last-bit effects; reported, and top-1
agreement asserted)."""
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
    for name in ('model_vit_test_w4a4', 'model_swin_test_w4a4'):
        g = torch.load(os.path.join(ROOT, 'tests', 'golden', name + '.pt'), weights_only=False)
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
            assert same_ranks and agree == 1.0
        dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
