"""Loader tail: CenterCropInKspace + max normalisation through the gather
kernels (csmri_shift_crop / csmri_plane_divide) vs the same steps as torch index ops around
the same library FFTs; kernel launches counted with the torch profiler."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from csmri_refinement_b200 import ops, rec_transforms as rt  # noqa: E402
from tools.gpu_time_aux import time_fn  # noqa: E402


def torch_chain(images, size):
    B, nx, ny = images.shape
    x = torch.stack([images, torch.zeros_like(images)], dim=1)
    x = torch.roll(x, shifts=(-(nx // 2), -(ny // 2)), dims=(2, 3))
    k = ops.fft2_planar(x.contiguous())
    k = torch.roll(k, shifts=(nx // 2, ny // 2), dims=(2, 3))
    r = size // 2
    crop = k[:, :, nx // 2 - r:nx // 2 + r, ny // 2 - r:ny // 2 + r]
    crop = torch.roll(crop, shifts=(-r, -r), dims=(2, 3))
    y = ops.fft2_planar(crop.contiguous(), inverse=True)
    y = torch.roll(y, shifts=(r, r), dims=(2, 3))
    m = rt.magnitude(y.contiguous())[:, 0]
    return m / m.abs().amax(dim=(1, 2), keepdim=True)


def launches(fn):
    fn()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    return len(ev), [(e.name.split('(')[0][-40:], round(e.device_time, 1)) for e in ev]


def main():
    out = []
    for B, n, size in ((256, 256, 256), (64, 512, 256), (164, 320, 256)):
        im = torch.rand(B, n, n, device='cuda')
        a, b = torch_chain(im, size), rt.crop_and_normalize(im, size)
        r = {'B': B, 'N': n, 'crop': size, 'bit_identical': bool(torch.equal(a, b))}
        r['torch_chain_us'] = round(time_fn(lambda: torch_chain(im, size)), 1)
        r['gather_us'] = round(time_fn(lambda: rt.crop_and_normalize(im, size)), 1)
        r['torch_chain_launches'], _ = launches(lambda: torch_chain(im, size))
        r['gather_launches'], r['gather_kernels'] = launches(lambda: rt.crop_and_normalize(im, size))
        print(json.dumps(r), flush=True)
        out.append(r)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'r2_loader_tail.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
