"""RecNet D5C5 training step with and without programmatic dependent launch for the
tensor-core convolution kernels (csmri_set_tuning key 10: 0 plain, 1 wait before the first
global read, 2 wait after the weight staging)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from csmri_refinement_b200 import _lib  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    torch.cuda.set_device(dev)
    out = []
    for rep in range(2):
        for mode in (0, 1, 2):
            _lib.lib().csmri_set_tuning(10, mode)
            r = bench.recnet_train_bench(dev, 0, 1, steps=10, warmup=3)
            row = {'pdl': mode, 'rep': rep, 'ms_per_step': round(r['ms_per_step'], 3), 'loss': r.get('loss')}
            print(json.dumps(row), flush=True)
            out.append(row)
    _lib.lib().csmri_set_tuning(10, 0)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r2_conv_pdl.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
