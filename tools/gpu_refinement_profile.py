"""Per-kernel time of one adversarial training step of configs/2-refinement.json
(torch.profiler, CUDA activities).  REFINE_TF32=0 keeps the U-Net / discriminator /
VGG19 convolutions in fp32 (default 1 = torch's stock TF32 convolutions; the frozen
RecNet path is always fp32)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from csmri_refinement_b200 import harness, refinement_harness as rh  # noqa: E402

tf32 = os.environ.get('REFINE_TF32', '1') == '1'
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cuda.matmul.allow_tf32 = tf32
torch.backends.cudnn.benchmark = True
dev = torch.device('cuda:0')
conf = harness.load_config(harness.config_path('2-refinement.json'))
tr = rh.AdversarialTrainer(conf, dev, channels_last=os.environ.get('REFINE_CL', '0') == '1')
batch = harness.synthetic_batch(conf, int(conf.batch_size), dev, seed=1)
for _ in range(3):
    tr.step(batch)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    tr.step(batch)
b.record()
torch.cuda.synchronize()
print('tf32=%d channels_last=%s  step %.2f ms (CUDA events, 3 steps)' % (tf32, os.environ.get('REFINE_CL', '0'), a.elapsed_time(b) / 3))
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
print('total kernel time per step %.2f ms, %d launches' % (tot / 1e3, sum(r[1] for r in rows_)))
for t, c, k in rows_[:28]:
    print('%9.1f us %5.1f%% x%-4d %s' % (t, 100 * t / tot, c, k[:120]))
