"""RecNet D5C5 training-step time under different conv memory formats (context probe)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from csmri_refinement_b200 import parallel, recnet, undersampling  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = torch.device('cuda:0')
B, n = 32, 256
img = torch.rand(B, n, n, device=dev)
rows = undersampling.cartesian_rows((B, n, n), 4, 8, False, np.random.RandomState(0))
batch = undersampling.undersample(img, rows)
for fmt in ('nchw', 'channels_last'):
    for graph in (True,):
        torch.manual_seed(0)
        model = recnet.construct_model({'num_blocks': 5, 'num_convs': 5, 'num_filters': 32}).to(dev)
        if fmt == 'channels_last':
            model = model.to(memory_format=torch.channels_last)
        tr = parallel.ShardedTrainer(model, lr=2e-4, cuda_graph=graph, assume_row_constant=True)
        for _ in range(3):
            tr.step(batch)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(6):
            loss = tr.step(batch)
        b.record()
        torch.cuda.synchronize()
        print(fmt, 'graph' if graph else 'eager', 'ms/step %.2f' % (a.elapsed_time(b) / 6),
              'loss %.6f' % float(loss), flush=True)
# forward-only split: convs vs DC
torch.manual_seed(0)
model = recnet.construct_model({'num_blocks': 5, 'num_convs': 5, 'num_filters': 32}).to(dev)
with torch.no_grad():
    for _ in range(2):
        model(batch['inp'], batch['kspace'], batch['mask'])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        model(batch['inp'], batch['kspace'], batch['mask'])
    b.record()
    torch.cuda.synchronize()
    print('forward only ms %.2f' % (a.elapsed_time(b) / 5))
