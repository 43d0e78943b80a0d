"""CPU-side tests: the C-ABI library builds, loads and exports every symbol
include/csmri_dc.h declares; host logic (mask line selection, RNG order,
RecNet mirror) matches the reference-generated fixtures; the product path
refuses to run without CUDA."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import dc_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='session')
def built_lib():
    from csmri_refinement_b200 import _lib
    _lib.build()
    return _lib


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, 'include', 'csmri_dc.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(csmri_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 12
    handle = ctypes.CDLL(built_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), name
    assert declared == set(built_lib.ABI_SYMBOLS)
    assert built_lib.lib().csmri_version() >= 100


def test_abi_rejects_bad_arguments_without_touching_the_gpu(built_lib):
    lib = built_lib.lib()
    assert lib.csmri_dc_workspace_bytes(4, 256, 256) == 4 * 2 * 256 * 256 * 4
    assert lib.csmri_dc_workspace_bytes(0, 256, 256) == 0
    # unsupported size -> CSMRI_E_SHAPE (-1), never a fallback
    rc = lib.csmri_fft2(8, 8, 1, 96, 96, 0, 8, None)
    assert rc == -1
    assert b'unsupported slice size' in lib.csmri_last_error()
    rc = lib.csmri_dc_forward_cartesian(None, None, None, None, None, 1, 256, 256, None)
    assert rc == -2 and b'NULL' in lib.csmri_last_error()
    rc = lib.csmri_dc_forward_cartesian(6, None, 8, None, 16, 1, 256, 256, None)
    assert rc == -3 and b'aligned' in lib.csmri_last_error()
    rc = lib.csmri_dc_adjoint_cartesian(8, 8, 8, 1, 256, 256, None)
    assert rc == -5                                   # aliasing
    with pytest.raises(RuntimeError, match='unsupported slice size'):
        built_lib.check(lib.csmri_fft2(8, 8, 1, 48, 48, 0, 8, None))


def test_header_is_plain_c_and_usable_without_python(built_lib, tmp_path):
    """include/csmri_dc.h compiles as pedantic C99 and a dlopen client gets the
    documented error codes (tests/abi_client.c)."""
    exe = str(tmp_path / 'abi_client')
    subprocess.run(['gcc', '-std=c99', '-pedantic', '-Wall', '-Wextra', '-Werror',
                    '-I', os.path.join(ROOT, 'include'), '-o', exe,
                    os.path.join(ROOT, 'tests', 'abi_client.c'), '-ldl'], check=True)
    out = subprocess.run([exe, built_lib.LIB_PATH], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert 'abi ok' in out.stdout


def test_sass_is_blackwell_native(built_lib):
    """The hot kernel uses the packed fp32 pipe (FFMA2/FADD2, sm_100 only)."""
    sass = subprocess.run(['cuobjdump', '-sass', built_lib.LIB_PATH], capture_output=True,
                          text=True, check=True).stdout
    assert 'sm_100a' in sass or 'SM100' in sass.upper() or 'arch = sm_100' in sass
    assert 'FFMA2' in sass and 'FADD2' in sass


def test_host_emulation_of_the_thread_choreography(tmp_path):
    """csrc/host_emulation.cu runs the exact LineFFT templates the kernels use,
    thread by thread on the host, against a naive DFT."""
    src = os.path.join(ROOT, 'csmri-refinement_b200', 'csrc', 'host_emulation.cu')
    exe = str(tmp_path / 'hostemu')
    subprocess.check_call(['nvcc', '-std=c++17', '-O1', '-Wno-deprecated-gpu-targets',
                           '-o', exe, src])
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert 'worst=' in res.stdout


def test_cartesian_rows_bit_exact_vs_reference(golden_dir):
    from csmri_refinement_b200 import undersampling as us
    g = np.load(os.path.join(golden_dir, 'masks.npz'))
    for key in g.files:
        if not key.startswith('rows_'):
            continue
        n = int(key.split('_')[1][1:])
        acc = int(key.split('_')[2][3:])
        rows = us.cartesian_rows((3, n, n), acc, 8, False, np.random.RandomState(0))
        assert rows.dtype == np.uint8
        assert np.array_equal(rows, g[key]), key
    m = us.cartesian_mask((2, 32, 32), 4, 8, False, np.random.RandomState(0))
    assert np.array_equal(m, g['full_n32_acc4'])
    m = us.cartesian_mask((2, 32, 32), 4, 8, True, np.random.RandomState(0))
    assert np.array_equal(m, g['full_n32_acc4_centred'])
    m = us.cartesian_mask((2, 64, 64), 3.5, rng=np.random.RandomState(7))
    assert np.array_equal(m, g['full_n64_acc3p5_s10'])


def test_undersample_transform_rng_order_matches_reference():
    """Fixed masks come from RandomState(0) at construction and the same rng
    then feeds the (zero) noise draws (myImageTransformations.py:1199-1207,1227)."""
    from csmri_refinement_b200 import undersampling as us
    n = 32
    tr = us.Undersample('varden', (1, n, n), 4, fixed_mask=True, num_fixed_masks=2)
    rng = np.random.RandomState(0)
    want = [orc.cartesian_mask((1, n, n), 4, 8, False, rng)[:, :, 0] for _ in range(2)]
    rows = tr.next_rows(3)
    assert np.array_equal(rows[0], want[0][0]) and np.array_equal(rows[1], want[1][0])
    assert np.array_equal(rows[2], want[0][0])          # cycles
    for _ in range(3):
        rng.normal(0, 1, (1, n, n))
        rng.normal(0, 1, (1, n, n))
    assert tr.rng.normal() == rng.normal()              # streams in step
    # per-sample masks from the global np.random, in the reference's order
    np.random.seed(3)
    tr2 = us.Undersample('varden', (1, n, n), 4)
    got = tr2.next_rows(2)
    np.random.seed(3)
    for i in range(2):
        m = orc.cartesian_mask((1, n, n), 4, 8, False, np.random)
        orc.undersample(np.zeros((1, n, n)), m, rng=np.random)
        assert np.array_equal(got[i], m[0, :, 0])


def test_recnet_mirror_matches_reference_on_cpu_with_oracle_dc(golden_dir):
    """Conv structure, state_dict keys and init RNG order of the mirror equal
    the reference's (DC swapped for the oracle, as the golden was made)."""
    from csmri_refinement_b200 import recnet
    g = np.load(os.path.join(golden_dir, 'recnet_tiny.npz'))
    torch.manual_seed(0)
    net = recnet.construct_model({'num_blocks': 2, 'num_convs': 3, 'num_filters': 4},
                                 dc_factory=orc.OracleDataConsistencyInKspace)
    sd = net.state_dict()
    ref_keys = {k[2:] for k in g.files if k.startswith('w:')}
    assert set(sd) == ref_keys
    for k in ref_keys:                                 # same seed -> same weights
        assert np.array_equal(sd[k].numpy(), g['w:' + k]), k
    inp, ksp, msk, tgt = (torch.from_numpy(g[k]) for k in ('inp', 'kspace', 'mask', 'target'))
    out = net(inp, ksp, msk)
    assert orc.rel_l2(out.detach().numpy(), g['out']) < 1e-6
    loss = torch.nn.functional.mse_loss(out, tgt)
    loss.backward()
    scale = max(np.linalg.norm(g[k]) for k in g.files if k.startswith('g:'))
    for name, p in net.named_parameters():
        want = g['g:' + name]
        if np.linalg.norm(want) < 1e-6 * scale:
            assert np.linalg.norm(p.grad.numpy() - want) < 1e-6 * scale, name
        else:
            assert orc.rel_l2(p.grad.numpy(), want) < 1e-5, name
    net2 = recnet.RecNet(2, 3, 4, use_refinement=True, return_intermediate_recs=True,
                         skip_final_dc=True, dc_factory=orc.OracleDataConsistencyInKspace)
    net2.load_state_dict(sd)
    o2 = net2(inp, ksp, msk)
    assert orc.rel_l2(o2['pred'].detach().numpy(), g['out_refine_pred']) < 1e-6
    assert len(o2['reconstructions']) == int(g['n_refine_recs'])
    assert list(net.forward.__code__.co_varnames[:4]) == ['self', 'inp', 'kspace', 'mask']
    assert not any(True for _ in net.buffers())        # DC adds no state


def test_product_path_refuses_cpu_tensors():
    from csmri_refinement_b200 import myfft
    x = torch.zeros(1, 2, 32, 32)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        myfft.DataConsistencyInKspace().perform(x, x, x)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'csmri-refinement_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_conv_module_is_a_plain_conv2d_on_cpu():
    """csmri_refinement_b200.conv.Conv2d keeps nn.Conv2d's parameters, keys and
    values; on tensors the kernels do not cover (CPU here) it is torch's own path
    plus the activation it was asked to apply."""
    from csmri_refinement_b200 import conv
    torch.manual_seed(0)
    m = conv.Conv2d(32, 32, 3, padding=1)
    ref = torch.nn.Conv2d(32, 32, 3, padding=1)
    assert list(m.state_dict().keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(m.state_dict())
    x = torch.randn(2, 32, 16, 32, requires_grad=True)
    assert torch.equal(m(x), ref(x))
    m.fused_slope = 0.01
    out = m(x)
    assert torch.equal(out, torch.nn.functional.leaky_relu(ref(x), 0.01))
    out.sum().backward()
    assert m.weight.grad is not None and x.grad is not None
    with pytest.raises(RuntimeError):
        conv.conv3x3_wgrad(x.detach(), out.detach(), 1)     # no CPU implementation
