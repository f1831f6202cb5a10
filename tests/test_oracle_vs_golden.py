"""CPU tests (-m "not gpu"): the oracle restatement (oracle/daft_exprt_oracle.py) against the golden vectors produced by
the real reference module (tests/golden/make_golden.py), and against the live reference when /root/reference exists."""
import numpy as np
import pytest
import torch

import daft_exprt_oracle as oracle
import reference_shims
from daft_exprt_b200 import synthetic
from helpers import (GOLDEN_CASES, case_inference_inputs, case_inputs, fake_stats, grad_projection, load_golden,
                     scale_rel_err, targets_of)

TOL = 2e-5   # fp32 CPU restatement vs fp32 CPU reference: same math, different op order


def oracle_state(n_ids, requires_grad=False, nudge=False):
    from daft_exprt_b200.model import reference_state_shapes
    sd = synthetic.synthetic_state_dict(reference_state_shapes(n_speakers=n_ids + 1), 1234)
    if nudge:
        sd = synthetic.nudge_for_inference(sd)
    if requires_grad:
        sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    return sd


@pytest.mark.parametrize('case', GOLDEN_CASES)
def test_forward_loss_grads_match_reference_golden(case):
    fx = load_golden(case)
    inputs, n_ids = case_inputs(fx)
    hp = oracle.OracleHParams(n_speakers=n_ids + 1)
    sd = oracle_state(n_ids, requires_grad=True)
    out = oracle.forward(sd, hp, inputs)
    spk, film, enc, dec, align = out
    assert torch.equal(dec[1], torch.from_numpy(fx['output_lengths']))            # integer outputs: bit-exact
    assert tuple(align.shape) == fx['alignments'].shape                           # T_max = max(cumsum) bit-exact
    for name, got in (('speaker_preds', spk), ('encoder_film', film[1]), ('prosody_pred_film', film[2]),
                      ('decoder_film', film[3]), ('duration_preds', enc[0]), ('energy_preds', enc[1]),
                      ('pitch_preds', enc[2]), ('mel_spec_preds', dec[0])):
        assert scale_rel_err(got.detach(), fx[name]) < TOL, name
    assert scale_rel_err(align.detach(), fx['alignments'].astype(np.float32)) < 1e-3  # stored as fp16
    total, terms = oracle.loss(hp, out, targets_of(inputs), int(fx['meta_iteration']))
    assert abs(total.item() - float(fx['loss_total'])) < TOL * abs(float(fx['loss_total']))
    for k, v in terms.items():
        ref = float(fx['loss_' + k])
        assert abs(float(v.detach()) - ref) <= TOL * max(abs(ref), 1e-6), k
    total.backward()
    names = [str(n) for n in fx['grad_names']]
    for i, n in enumerate(names):
        g = sd[n].grad
        assert g is not None, n
        ref_norm = float(fx['grad_norms'][i])
        assert abs(g.double().norm().item() - ref_norm) <= 1e-4 * max(ref_norm, 1e-8), n
        assert abs(grad_projection(n, g) - float(fx['grad_projs'][i])) <= 2e-4 * max(4 * ref_norm, 1e-8), n
        if 'grad:' + n in fx.files:
            assert scale_rel_err(g, fx['grad:' + n]) < 5e-4, n   # fp32 reduction-order noise over B*T rows


@pytest.mark.parametrize('case', GOLDEN_CASES)
@pytest.mark.parametrize('transform', ['add', 'multiply'])
def test_inference_matches_reference_golden(case, transform):
    fx = load_golden(case)
    n_ids = int(fx['meta_n_speaker_ids'])
    hp = oracle.OracleHParams(n_speakers=n_ids + 1, stats=fake_stats(n_ids))
    sd = oracle_state(n_ids, nudge=True)
    with torch.no_grad():
        enc, dec, w = oracle.inference(sd, hp, case_inference_inputs(fx, transform), transform)
    assert torch.equal(enc[1], torch.from_numpy(fx[f'inf_{transform}_durations_int']))   # bit-exact ints
    assert torch.equal(dec[1], torch.from_numpy(fx[f'inf_{transform}_output_lengths']))
    for name, got in (('duration_preds', enc[0]), ('energy_preds', enc[2]), ('pitch_preds', enc[3]), ('mel_spec_preds', dec[0])):
        assert scale_rel_err(got, fx[f'inf_{transform}_{name}']) < 5e-5, name


def test_int_durations_known_answers():
    kat = np.load(__import__('os').path.join(__import__('helpers').GOLDEN_DIR, 'int_durations_kat.npz'))
    hp = oracle.OracleHParams()
    _, got = oracle.get_int_durations(torch.from_numpy(kat['durations']), hp)
    assert torch.equal(got, torch.from_numpy(kat['durations_int']))


@pytest.mark.skipif(not reference_shims.reference_available(), reason='/root/reference not present (GPU box)')
def test_oracle_matches_live_reference_all_grads():
    """Authoring-container only: run the unmodified reference and compare EVERY gradient tensor in full."""
    ref_model, ref_loss, _, _ = reference_shims.install()
    n_ids = 3
    hp_ref = reference_shims.make_reference_hparams(n_ids)
    model = ref_model.DaftExprt(hp_ref)
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(synthetic.synthetic_state_dict(shapes, 77))
    model.eval()
    inputs = synthetic.make_batch(3, 23, 90, n_ids, seed=5)
    total_ref, _ = ref_loss.DaftExprtLoss('cpu', hp_ref)(model(inputs), targets_of(inputs), 100)
    total_ref.backward()
    sd = {k: v.clone().requires_grad_(True) for k, v in synthetic.synthetic_state_dict(shapes, 77).items()}
    hp = oracle.OracleHParams(n_speakers=n_ids + 1)
    total, _ = oracle.loss(hp, oracle.forward(sd, hp, inputs), targets_of(inputs), 100)
    total.backward()
    assert abs(total.item() - total_ref.item()) < 1e-5 * abs(total_ref.item())
    for n, p in model.named_parameters():
        assert scale_rel_err(sd[n].grad, p.grad) < 2e-4, n
