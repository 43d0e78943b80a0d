"""Workload for the f-3 question: do the 2-channel tensors around the DC layers of a
RecNet training step round-trip HBM at all on a B200 (126 MB L2)?

Run under ncu WITHOUT cache flushing between kernels:

  ncu --cache-control none --clock-control none --metrics \
      dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum \
      -k regex:'dc_strip|conv3x3_thin' -c 60 --csv --log-file gpurun_out/r2_recnet_l2.csv \
      python tools/gpu_recnet_l2_probe.py

One eager RecNet D5C5 step (batch 32, 256^2): the DC kernel's x / addend / out
are 16 MiB each and the addend is re-read by all five cascades."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import parallel, recnet, undersampling  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
B, n = int(os.environ.get('PROBE_B', '32')), 256
img = torch.rand(B, n, n, device=dev)
rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
batch = undersampling.undersample(img, rows)
torch.manual_seed(0)
model = recnet.construct_model({'num_blocks': 5, 'num_convs': 5, 'num_filters': 32}).to(dev)
tr = parallel.ShardedTrainer(model, lr=2e-4, cuda_graph=False, assume_row_constant=True)
for _ in range(2):
    tr.step(batch)
torch.cuda.synchronize()
print('probe done, loss %.5f' % float(tr.step(batch).item()))
