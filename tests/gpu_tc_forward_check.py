"""GPU helper (not a pytest file): calibration time and fake-quant inference throughput with the tensor-core inference
forward on/off (ADALOG_B200_TC_FORWARD), printed from bench.py's own JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for t in ('0', '1'):
    env = dict(os.environ, ADALOG_B200_TC_FORWARD=t)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--no-cpu-baseline'], env=env,
                         capture_output=True, text=True, timeout=600).stdout.strip().splitlines()[-1]
    d = json.loads(out)
    print(f'TC_FORWARD={t}: {d["ms_per_step"] / 1e3:.2f} s per calibration, fake-quant forward {d["fakequant_img_per_s"]:.0f} img/s')
