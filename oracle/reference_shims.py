"""Harness-side shims that let the UNMODIFIED reference (`/root/reference/src/daft_exprt`) be imported and
run on CPU inside the authoring container.  TEST INFRASTRUCTURE ONLY: used by `tests/golden/make_golden.py`
and by the `not gpu` tests that cross-check the oracle when `/root/reference` is present.  Nothing in the
product path, `bench.py` or `smoke()` imports this file, and `/root/reference` does not exist on the GPU box.

Shims:
  1. stub modules for librosa / matplotlib (imported by extract_features.py:10,16 and utils.py:7; not on
     the arithmetic path);
  2. `HyperParams.update_mfa_paths` -> no-op (hparams.py:219-230 asserts MFA model files exist on disk);
  3. `torch.Tensor.cuda(device)` with a CPU device returns `self` (model.py:22,139,651,810,913, loss.py:60
     call `.cuda(x.device)` which raises on CPU tensors).
"""
import os
import sys
import types

REFERENCE_SRC = os.environ.get('DAFT_EXPRT_REFERENCE_SRC', '/root/reference/src')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, 'daft_exprt'))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install():
    """Install the three shims and return the reference's (model, loss, hparams, extract_features) modules."""
    import torch
    if not reference_available():
        raise RuntimeError(f'reference not found under {REFERENCE_SRC}')
    for name in ('librosa', 'matplotlib'):
        try:
            __import__(name)
        except Exception:
            pass
    if 'librosa' not in sys.modules:
        lib = _stub('librosa')
        lib.filters = _stub('librosa.filters', mel=lambda *a, **k: None)
        lib.util = _stub('librosa.util')
        lib.core = _stub('librosa.core')
    if 'matplotlib' not in sys.modules:
        mpl = _stub('matplotlib', use=lambda *a, **k: None)
        mpl.pyplot = _stub('matplotlib.pyplot')
    for name in ('tgt', 'inflect', 'unidecode', 'tensorboard'):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)

    if not getattr(torch.Tensor.cuda, '_dx_shim', False):
        _orig_cuda = torch.Tensor.cuda

        def _cuda(self, device=None, *args, **kwargs):
            if isinstance(device, torch.device) and device.type == 'cpu':
                return self
            if isinstance(device, str) and device == 'cpu':
                return self
            return _orig_cuda(self, device, *args, **kwargs)
        _cuda._dx_shim = True
        torch.Tensor.cuda = _cuda
        _orig_mod_cuda = torch.nn.Module.cuda

        def _mod_cuda(self, device=None):
            if device == 'cpu' or (isinstance(device, torch.device) and device.type == 'cpu'):
                return self
            return _orig_mod_cuda(self, device)
        torch.nn.Module.cuda = _mod_cuda

    import daft_exprt.hparams as ref_hparams
    import daft_exprt.model as ref_model
    import daft_exprt.loss as ref_loss
    import daft_exprt.extract_features as ref_feats
    return ref_model, ref_loss, ref_hparams, ref_feats


def make_reference_hparams(n_speakers_ids, stats=None, **overrides):
    """Build the reference's own HyperParams object without MFA files on disk (shim 2)."""
    _, _, ref_hparams, _ = install()

    class _HP(ref_hparams.HyperParams):
        def update_mfa_paths(self):
            return None

    speakers = [f'spk{i}' for i in range(n_speakers_ids)]
    kwargs = dict(training_files='unused', validation_files='unused', output_directory='unused',
                  language='english', speakers=speakers)
    kwargs.update(overrides)
    hp = _HP(verbose=False, **kwargs)
    if stats is not None:
        hp.stats = stats
    return hp
