"""adalog_b200: B200-native (sm_100a) implementation of AdaLog's FPCS calibration sweep and the
fake-quant forward it scores, behind the reference's Python API (quantizers/, quant_layers/,
utils/calibrator.py, utils/wrap_net.py).  See DESIGN.md."""
__version__ = '0.1.0'
