"""Per-kernel time of one eager RecNet D5C5 training step (torch.profiler, CUDA activities).
RECNET_TF32=1 profiles torch's stock TF32 convolutions instead of fp32."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from csmri_refinement_b200 import parallel, recnet, undersampling  # noqa: E402

tf32 = os.environ.get('RECNET_TF32', '0') == '1'
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device('cuda:0')
B, n = 32, 256
img = torch.rand(B, n, n, device=dev)
rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
batch = undersampling.undersample(img, rows)
torch.manual_seed(0)
model = recnet.construct_model({'num_blocks': 5, 'num_convs': 5, 'num_filters': 32}).to(dev)
tr = parallel.ShardedTrainer(model, lr=2e-4, cuda_graph=False, assume_row_constant=True)
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        tr.step(batch)
    torch.cuda.synchronize()
rows_ = []
for e in prof.key_averages():
    t = getattr(e, 'device_time_total', None)
    if t is None:
        t = e.cuda_time_total
    rows_.append((t / 2.0, e.count // 2, e.key))
rows_.sort(reverse=True)
tot = sum(r[0] for r in rows_)
print('tf32=%d  total kernel time per step %.2f ms in %d launches' % (tf32, tot / 1e3, sum(r[1] for r in rows_)))
for t, c, k in rows_[:40]:
    print('%9.1f us %5.1f%% x%-4d %s' % (t, 100 * t / tot, c, k[:110]))
