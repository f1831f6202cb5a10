"""Generate the golden fixtures under tests/golden/ by executing the UNMODIFIED reference module.

Run in the authoring container only (needs /root/reference):   python tests/golden/make_golden.py
It imports `daft_exprt.model.DaftExprt` / `daft_exprt.loss.DaftExprtLoss` through the shims of
`oracle/reference_shims.py`, loads name-keyed synthetic weights (`synthetic_state_dict`), runs the seeded synthetic
batches of `synthetic.make_batch` in eval mode on CPU fp32 and stores:
  * every output of `forward` (mel, alignments, predictions, FiLM tensors, speaker logits),
  * the 7 loss terms + total,
  * per-parameter gradient L2 norms and a fixed random projection of every gradient, full grads of small tensors,
  * `inference()` outputs (incl. the integer durations) for both pitch transforms,
  * known-answer vectors for `duration_to_integer` / `get_int_durations`.
Weights and inputs are NOT stored: they are regenerated bit-identically from the seeds recorded in the fixture.
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import reference_shims  # noqa: E402
from daft_exprt_b200 import synthetic  # noqa: E402

CASES = {
    # name: (n_speaker_ids, B, L, T, batch_seed)
    'c1_1spk_b4': (1, 4, 50, 250, 11),          # BASELINE.json configs[0]
    'c2s_11spk_b6': (11, 6, 120, 520, 12),      # reduced configs[1] (same hparams, smaller batch)
    'ragged_11spk_b5': (11, 5, 37, 131, 13),    # odd sizes, not multiples of any tile
}
WEIGHT_SEED = 1234


def grad_projection(name, g):
    rng = np.random.RandomState(zlib.crc32(('proj:' + name).encode()) & 0x7fffffff)
    v = rng.randn(g.numel()).astype(np.float64)
    return float(np.dot(g.detach().double().numpy().ravel(), v))


def fake_stats(n_ids):
    return {f'spk {i}': {'pitch': {'mean': 5.0 + 0.05 * i, 'std': 0.25 + 0.01 * i}} for i in range(n_ids)}


def main():
    ref_model, ref_loss, _, ref_feats = reference_shims.install()
    torch.set_num_threads(8)
    for case, (n_ids, B, L, T, seed) in CASES.items():
        hp = reference_shims.make_reference_hparams(n_ids, stats=fake_stats(n_ids))
        torch.manual_seed(0)
        model = ref_model.DaftExprt(hp)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        model.load_state_dict(synthetic.synthetic_state_dict(shapes, WEIGHT_SEED))
        model.eval()
        criterion = ref_loss.DaftExprtLoss('cpu', hp)

        inputs = synthetic.make_batch(B, L, T, n_ids, seed=seed)
        outputs = model(inputs)
        targets = (inputs[1], inputs[3], inputs[4], inputs[8], inputs[10])
        iteration = 2500
        total, terms = criterion(outputs, targets, iteration)
        model.zero_grad()
        total.backward()

        spk, film, enc, dec, align = outputs
        fx = {
            'meta_n_speaker_ids': np.int64(n_ids), 'meta_B': np.int64(B), 'meta_L': np.int64(L), 'meta_T': np.int64(T),
            'meta_batch_seed': np.int64(seed), 'meta_weight_seed': np.int64(WEIGHT_SEED), 'meta_iteration': np.int64(iteration),
            'speaker_preds': spk.detach().numpy(),
            'encoder_film': film[1].detach().numpy(), 'prosody_pred_film': film[2].detach().numpy(),
            'decoder_film': film[3].detach().numpy(),
            'duration_preds': enc[0].detach().numpy(), 'energy_preds': enc[1].detach().numpy(),
            'pitch_preds': enc[2].detach().numpy(),
            'mel_spec_preds': dec[0].detach().numpy(), 'output_lengths': dec[1].numpy(),
            'alignments': align.detach().numpy().astype(np.float16),  # (B, L, T) in [0,1]; fp16 keeps the fixture small
            'loss_total': np.float64(total.item()),
        }
        for k, v in terms.items():
            fx['loss_' + k] = np.float64(v)
        names = [n for n, _ in model.named_parameters()]
        fx['grad_names'] = np.array(names)
        fx['grad_norms'] = np.array([p.grad.double().norm().item() for _, p in model.named_parameters()])
        fx['grad_projs'] = np.array([grad_projection(n, p.grad) for n, p in model.named_parameters()])
        for n, p in model.named_parameters():
            if p.numel() <= 1024:
                fx['grad:' + n] = p.grad.numpy()

        # inference (model.py:866-923), both pitch transforms.  Nudge the duration bias so that predicted
        # durations land in a realistic range.
        with torch.no_grad():
            model.load_state_dict(synthetic.nudge_for_inference(synthetic.synthetic_state_dict(shapes, WEIGHT_SEED)))
            inf_inputs = synthetic.make_inference_batch(B, L, min(T, 300), n_ids, seed=seed)
            inf_inputs = list(inf_inputs)
            rng = np.random.RandomState(seed)
            inf_inputs[1] = torch.from_numpy((0.8 + 0.6 * rng.rand(B, L)).astype(np.float32))   # duration factors
            inf_inputs[2] = torch.from_numpy((0.9 + 0.2 * rng.rand(B, L)).astype(np.float32))   # energy factors
            fx['inf_dur_factors'] = inf_inputs[1].numpy()
            fx['inf_energy_factors'] = inf_inputs[2].numpy()
            for transform in ('add', 'multiply'):
                pf = (20.0 * rng.randn(B, L)).astype(np.float32) if transform == 'add' else (0.5 * rng.randn(B, L)).astype(np.float32)
                inf_inputs[3] = torch.from_numpy(pf)
                fx[f'inf_{transform}_pitch_factors'] = pf
                e, d, w = model.inference(tuple(t.clone() for t in inf_inputs), transform, hp)
                fx[f'inf_{transform}_duration_preds'] = e[0].numpy()
                fx[f'inf_{transform}_durations_int'] = e[1].numpy()
                fx[f'inf_{transform}_energy_preds'] = e[2].numpy()
                fx[f'inf_{transform}_pitch_preds'] = e[3].numpy()
                fx[f'inf_{transform}_mel_spec_preds'] = d[0].numpy()
                fx[f'inf_{transform}_output_lengths'] = d[1].numpy()
        np.savez_compressed(os.path.join(HERE, case + '.npz'), **fx)
        print(case, 'loss', total.item(), 'T_out(add)', int(fx['inf_add_output_lengths'].max()))

    # known-answer vectors for duration_to_integer (extract_features.py:69-111) through get_int_durations
    hp = reference_shims.make_reference_hparams(1)
    model = ref_model.DaftExprt(hp)
    rng = np.random.RandomState(99)
    rows = []
    for _ in range(64):
        n = rng.randint(3, 40)
        d = np.zeros(40, np.float32)
        d[:n] = rng.choice([0.0, 0.01, 0.0232, 0.0233, 0.03, 0.05, 0.08, 0.11, 0.2], size=n).astype(np.float32) \
            + (rng.rand(n) * 0.02).astype(np.float32) * (rng.rand(n) < 0.7)
        rows.append(d)
    durs = torch.from_numpy(np.stack(rows))
    good_in, good_out = [], []
    for r in range(durs.shape[0]):
        try:
            fl, it = model.get_int_durations(durs[r:r + 1].clone(), hp)
            good_in.append(durs[r].numpy())
            good_out.append(it[0].numpy())
        except Exception as exc:  # reference raises IndexError when the total is < filter_length samples
            print('row', r, 'reference raised', type(exc).__name__)
    np.savez_compressed(os.path.join(HERE, 'int_durations_kat.npz'),
                        durations=np.stack(good_in), durations_int=np.stack(good_out))
    print('int_durations KAT rows:', len(good_in))


if __name__ == '__main__':
    main()
