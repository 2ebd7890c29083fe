"""Calibration driver with the reference's API (utils/calibrator.py): QuantCalibrator(model, loader).batching_quant_calib().

Same protocol per module -- capture raw input/output of the still-FP32 model over the calibration set, run
hyperparameter_searching(), reparam() where a preceding LayerNorm is attached, finally flip every module to
'quant_forward' -- but the captured tensors stay in HBM (no .cpu() round trip per batch, reference :14-28) where the
device sweeps consume them directly.  Under torch.distributed each rank captures its own shard of the loader.
"""
import torch

from ..quant_layers import MinMaxQuantConv2d, MinMaxQuantLinear, MinMaxQuantMatMul

try:
    from tqdm import tqdm
except Exception:  # noqa: BLE001
    tqdm = None


class QuantCalibrator:
    def __init__(self, model, calib_loader):
        self.model = model
        self.calib_loader = calib_loader
        self.progress = True

    # hooks keep device tensors (detached), reference :14-28
    def single_input_forward_hook(self, module, inp, outp):
        if module.tmp_input is None:
            module.tmp_input = []
        module.tmp_input.append(inp[0].detach())

    def double_input_forward_hook(self, module, inp, outp):
        if module.tmp_input is None:
            module.tmp_input = [[], []]
        module.tmp_input[0].append(inp[0].detach())
        module.tmp_input[1].append(inp[1].detach())

    def outp_forward_hook(self, module, inp, outp):
        if module.tmp_out is None:
            module.tmp_out = []
        module.tmp_out.append(outp.detach())

    def _pending(self):
        return [(n, m) for n, m in self.model.named_modules() if hasattr(m, 'calibrated') and not m.calibrated]

    def _capture(self, module, device):
        hooks = [module.register_forward_hook(self.outp_forward_hook)]
        if isinstance(module, (MinMaxQuantLinear, MinMaxQuantConv2d)):
            hooks.append(module.register_forward_hook(self.single_input_forward_hook))
        if isinstance(module, MinMaxQuantMatMul):
            hooks.append(module.register_forward_hook(self.double_input_forward_hook))
        with torch.no_grad():
            for inp, _ in self.calib_loader:
                self.model(inp.to(device))
        module.raw_out = torch.cat(module.tmp_out, dim=0)
        if isinstance(module, MinMaxQuantMatMul):
            module.raw_input = [torch.cat(t, dim=0) for t in module.tmp_input]
        else:
            module.raw_input = torch.cat(module.tmp_input, dim=0)
        for h in hooks:
            h.remove()
        module.tmp_input = module.tmp_out = None

    def batching_quant_calib(self):
        """reference calibrator.py:30-67"""
        device = next(self.model.parameters()).device
        pending = self._pending()
        bar = tqdm(total=len(pending)) if (tqdm is not None and self.progress) else None
        for name, module in self.model.named_modules():
            if not hasattr(module, 'calibrated') or module.calibrated:
                continue
            if bar is not None:
                bar.set_description(f"calibrating {name}")
            self._capture(module, device)
            with torch.no_grad():
                module.hyperparameter_searching()
                if hasattr(module, 'prev_layer') and module.prev_layer is not None:
                    if bar is not None:
                        bar.set_description(f"reparaming {name}")
                    module.reparam()
            if bar is not None:
                bar.update()
        if bar is not None:
            bar.close()
        for _, module in self.model.named_modules():
            if hasattr(module, 'mode'):
                module.mode = "quant_forward"
