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
    # per-kernel: the hot strip kernels are TMA-fed (UTMALDG tensor loads landing on
    # an mbarrier = SYNCS), the thin convolutions stage with cp.async (LDGSTS)
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import sass_histogram
    hist = sass_histogram.histogram(sass)
    hot = [k for k in hist if 'dc_strip_pipev_kernel' in k]
    assert hot, 'two-column strip kernel missing from the library'
    for k in hot:
        assert hist[k].get('UTMALDG', 0) >= 1, k
        assert hist[k].get('SYNCS', 0) >= 1, k
        assert hist[k].get('FFMA2', 0) + hist[k].get('FADD2', 0) + hist[k].get('FMUL2', 0) >= 300, k
        assert hist[k].get('HMMA', 0) == 0 and hist[k].get('UTCHMMA', 0) == 0   # no tensor cores: not a contraction
    assert any(hist[k].get('LDGSTS', 0) for k in hist if 'conv3x3' in k)
    # the 32 -> 32 convolutions (the one GEMM-shaped piece) run on the 5th-generation
    # tensor cores: UTCHMMA = tcgen05.mma, accumulators read back with LDTM, and the
    # weight gradient feeds its A operand through STTM and its B operand through TMA
    tc = [k for k in hist if 'conv3x3_tc_kernel' in k]
    wtc = [k for k in hist if 'conv3x3_wgrad_tc_kernel' in k]
    assert tc and wtc, 'tensor-core convolution kernels missing from the library'
    for k in tc + wtc:
        assert hist[k].get('UTCHMMA', 0) >= 20, (k, hist[k].get('UTCHMMA', 0))
        assert hist[k].get('LDTM', 0) >= 1, k
        assert hist[k].get('HMMA', 0) == 0, k           # no legacy warp-level MMA
    for k in wtc:
        assert hist[k].get('UTMALDG', 0) >= 1 and hist[k].get('STTM', 0) >= 1, k


def test_build_staleness_covers_every_source(built_lib):
    """editing any csrc/*.cuh (not only the three files once listed)
    must trigger a rebuild."""
    srcs = [os.path.basename(p) for p in built_lib.sources()]
    csrc = os.path.join(ROOT, 'csmri-refinement_b200', 'csrc')
    on_disk = [f for f in os.listdir(csrc) if f.endswith(('.cu', '.cuh'))]
    assert set(on_disk) <= set(srcs) and 'csmri_dc.h' in srcs
    assert not built_lib.is_stale()
    probe = os.path.join(csrc, 'dc_pipev.cuh')
    st = os.stat(probe)
    try:
        os.utime(probe, (st.st_atime, os.path.getmtime(built_lib.LIB_PATH) + 5))
        assert built_lib.is_stale()
    finally:
        os.utime(probe, (st.st_atime, st.st_mtime))
    assert not built_lib.is_stale()


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


def test_configuration_mirror_reads_the_shipped_configs_unchanged(golden_dir):
    """configs/1-recnet.json and 2-refinement.json are byte-identical copies of
    the reference's files, parse through the Configuration mirror exactly as
    through utils/config.py:212-250, and construct_model on the first builds the
    reference's network: same keys, same shapes, SAME initial weights under the
    config's own seed (training/runner.py:18-21, models/recnet.py:20-26)."""
    import hashlib
    from csmri_refinement_b200 import harness
    from csmri_refinement_b200.config import Configuration
    g = np.load(os.path.join(golden_dir, 'configs.npz'))
    for name in ('1-recnet.json', '2-refinement.json'):
        path = harness.config_path(name)
        with open(path, 'rb') as f:
            data = f.read()
        assert hashlib.sha256(data).hexdigest() == str(g['sha256:' + name])
        ref = os.path.join('/root/reference/configs', name)
        if os.path.exists(ref):                        # build container only
            assert open(ref, 'rb').read() == data
    conf = harness.load_config(harness.config_path('1-recnet.json'))
    assert conf.seed == int(g['recnet1:seed']) and 'seed' not in conf.__dict__
    assert conf.file.endswith('1-recnet.json')
    assert isinstance(conf.model, dict) and conf.batch_size == int(g['recnet1:batch_size'])
    assert conf.get_attr('nonexistent', default=7) == 7
    assert conf.get_attr('nonexistent', alternative='batch_size') == 20
    with pytest.raises(ValueError):
        conf.get_attr('nonexistent', alternative='also_missing')
    harness.set_random_seeds(conf.seed)
    net = harness.build_recnet(conf)
    assert sum(p.numel() for p in net.parameters()) == int(g['recnet1:num_params']) == 31302
    assert len(net.dc_layers) == int(g['recnet1:num_dc'])
    assert list(net.state_dict().keys()) == [str(k) for k in g['recnet1:keys']]
    for k, v in net.state_dict().items():
        assert np.array_equal(v.numpy(), g['recnet1:w:' + k]), k
    assert harness.adam_args(conf.optimizer, conf) == {'lr': float(g['recnet1:lr']),
                                                      'betas': (0.9, 0.999)}
    assert harness.undersampling_args(conf) == {'acc': int(g['recnet1:acc']), 'variable': False}
    conf2 = harness.load_config(harness.config_path('2-refinement.json'))
    assert conf2.seed == int(g['refine2:seed'])
    assert sorted(k for k in conf2.__dict__ if not k.startswith('_')) == \
        [str(k) for k in g['refine2:top_keys']]
    gen = Configuration.from_dict(conf2.generator_model, conf2)
    assert gen.mode == str(g['refine2:gen_mode']) and gen.seed == conf2.seed
    disc = Configuration.from_dict(conf2.discriminator_model, conf2)
    assert disc.has_attr('name') == bool(g['refine2:disc_has_name'])   # it has none
    assert harness.adam_args(conf2.generator_optimizer, conf2)['betas'] == (0.5, 0.999)
    # --conf key=value overrides and the include mechanisms
    conf.update({'batch_size': '8', 'seed': '3', 'tags': '[a, 1, 2.5]', 'flag': 'True'})
    assert conf.batch_size == 8 and conf.seed == 3 and conf.tags == ['a', 1, 2.5] and conf.flag is True


def test_configuration_includes(tmp_path):
    """`#include` and the top-level `include` section behave as in
    utils/config.py:10-19,236-250 (checked against the reference in the build
    container: identical __dict__ for this input)."""
    from csmri_refinement_b200.config import Configuration
    (tmp_path / 'base.json').write_text('{"seed": 4, "a": 1, "model": {"x": 1, "y": 2}}')
    (tmp_path / 'm.json').write_text('{"x": 10, "z": 30}')
    (tmp_path / 'top.json').write_text(
        '{"#include": "base.json", "a": 2, "include": {"model": "m.json", "opt": "m.json"}}')
    conf = Configuration.from_json(str(tmp_path / 'top.json'))
    assert conf.seed == 4 and conf.a == 2 and not conf.has_attr('include')
    strip = lambda d: {k: v for k, v in d.items() if not k.startswith('_')}   # noqa: E731
    assert strip(conf.model) == {'x': 1, 'z': 30, 'y': 2}   # own keys win over the included file
    assert strip(conf.opt) == {'x': 10, 'z': 30}
    assert conf.to_param_dict(['a'], ['opt', 'missing'], {'a': 'alpha'})['alpha'] == 2
    assert conf.to_param_dict([], {'missing': 5}) == {'missing': 5}
    if os.path.exists('/root/reference/utils/config.py'):
        sys.path.insert(0, '/root/reference')
        try:
            from utils.config import Configuration as Ref
        finally:
            sys.path.remove('/root/reference')
        ref = Ref.from_json(str(tmp_path / 'top.json'))
        assert strip(ref.model) == strip(conf.model) and strip(ref.opt) == strip(conf.opt)
        assert ref.seed == conf.seed and ref.a == conf.a
