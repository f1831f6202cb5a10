"""CPU tests (-m "not gpu"): the C-ABI library builds, loads and exports every symbol include/daft_exprt_b200.h declares;
host-side logic (header parser, hparams surface, state-dict contract, synthetic batches).  No compute calls without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from daft_exprt_b200 import cabi, synthetic
from daft_exprt_b200.hparams import default_hparams
from daft_exprt_b200.model import DaftExprt, reference_state_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib_path():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(cabi.LIB_PATH)
    return cabi.LIB_PATH


def test_header_prototypes_parse():
    protos = cabi.parse_header()
    assert len(protos) >= 35
    assert protos['dx_last_error'][0] is ctypes.c_char_p
    assert protos['dx_conv_wgrad_workspace'][0] is ctypes.c_size_t
    assert protos['dx_attention_fwd'][1][-1] is ctypes.c_void_p and protos['dx_attention_fwd'][1][-2] is ctypes.c_uint64


def test_library_exports_every_declared_symbol(lib_path):
    lib = cabi.load(lib_path)
    for name in cabi.parse_header():
        assert hasattr(lib, name), name
    assert lib.dx_abi_version() == 1
    out = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if ' T ' in l}
    assert set(cabi.parse_header()) <= exported


def test_library_contains_sm100a_code_only(lib_path):
    out = subprocess.run(['cuobjdump', '-lelf', lib_path], capture_output=True, text=True).stdout
    archs = {l.split('.')[-2] for l in out.splitlines() if l.strip().endswith('.cubin')}
    assert archs == {'sm_100a'}, archs


def test_error_reporting_without_gpu(lib_path):
    lib = cabi.load(lib_path)
    assert lib.dx_set_gemm_backend(7) != 0
    assert b'backend' in lib.dx_last_error()
    if not torch.cuda.is_available():
        assert lib.dx_device_check() != 0          # loud failure, never a silent CPU path


def test_state_dict_contract_193_tensors():
    sh = reference_state_shapes(n_speakers=12)
    assert len(sh) == 193 and sum(int(np.prod(v)) for v in sh.values()) == 14727153
    assert sh['prosody_encoder.blocks.0.attention.multi_head_attention.in_proj_weight'] == (384, 128)
    assert sh['prosody_encoder.convs.4.conv.weight'] == (1024, 1024, 3)
    assert sh['prosody_encoder.post_multipliers'] == (2, 9)
    assert sh['prosody_predictor.projection.linear_layer.weight'] == (3, 256)
    assert sh['gaussian_upsampling.projection.0.linear_layer.weight'] == (1, 128)
    assert sh['frame_decoder.projection.linear_layer.weight'] == (80, 128)
    assert sh['speaker_classifier.classifier.5.linear_layer.weight'] == (11, 128)
    m = DaftExprt(default_hparams(n_speakers=2))
    assert m.state_dict()['speaker_classifier.classifier.5.linear_layer.weight'].shape == (1, 128)
    assert next(iter(dict(m.named_parameters()))) == 'prosody_encoder.post_multipliers'


def test_module_has_no_cpu_fallback():
    m = DaftExprt(default_hparams(n_speakers=4))
    with pytest.raises(RuntimeError):
        m(synthetic.make_batch(2, 9, 30, 3, seed=1))


def test_synthetic_batch_invariants():
    b = synthetic.make_batch(8, 50, 250, 11, seed=3)
    symbols, dur_f, dur_i, en, pi, in_len, fe, fp, mel, out_len, spk = b
    assert torch.equal(dur_i.sum(1), out_len) and int(out_len.max()) == 250 == mel.shape[2]      # data_loader.py:128
    assert (in_len[:-1] >= in_len[1:]).all() and int(in_len[0]) == 50                            # sorted by length
    assert float((dur_i == 0).float().mean()) > 0.1                                             # zero-duration symbols
    for r in range(8):
        assert (symbols[r, in_len[r]:] == 0).all() and (dur_i[r, in_len[r]:] == 0).all()
        assert (mel[r, :, out_len[r]:] == 0).all() and (fe[r, out_len[r]:] == 0).all()
    b2 = synthetic.make_batch(8, 50, 250, 11, seed=3)
    assert all(torch.equal(x, y) for x, y in zip(b, b2))
