"""CPU: the oracle restatements reproduce the golden vectors that oracle/gen_golden.py recorded
from the reference's own modules (integer outputs bit-exact, floats to fp32 round-off)."""
import os

import numpy as np
import pytest
import torch

from oracle import geometries, tte_oracle, voc_oracle
from parrot_tts_b200 import configs, synth


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", ["tte_b1_s16", "tte_b3_ragged", "tte_cfg1_s64", "tte_b4_s256_ragged", "tte_geom_l23_s41"])
def test_tte_oracle_matches_golden(golden_dir, name):
    g = _load(golden_dir, name)
    cfg = geometries.other_tte_config() if "geom" in name else configs.default_tte_config()
    sd = synth.synth_tte_state_dict(cfg, int(g["vocab"]), 0, int(g["n_speaker"]), seed=int(g["seed"]), dur_mode=str(g["dur_mode"]))
    batch = {"phones": torch.from_numpy(g["phones"]), "src_mask": torch.from_numpy(g["src_mask"]),
             "speaker": torch.from_numpy(g["speaker"])}
    ids, r = tte_oracle.parrot_infer(sd, batch)
    assert np.array_equal(r["dur"].numpy(), g["dur"])                       # integer: bit-exact
    assert np.array_equal(r["tgt_mask"].numpy(), g["tgt_mask"])
    assert [len(x) for x in ids] == g["ids_lens"].tolist()
    assert np.array_equal(np.concatenate([np.asarray(x) for x in ids]), g["ids_flat"])
    np.testing.assert_allclose(r["log_dur"].numpy(), g["log_dur"], rtol=0, atol=2e-5)
    stride = int(g["logits_row_stride"])
    np.testing.assert_allclose(r["logits"][:, ::stride].numpy(), g["logits_rows"], rtol=0, atol=5e-4)


def test_mask_off_by_one():
    # reference modules/data.py:18 `ids <= len` marks len+1 positions valid [probed in SURVEY 8a/a9]
    m = tte_oracle.get_mask_from_lengths([3, 5], 5)
    assert m.int().tolist() == [[1, 1, 1, 1, 0], [1, 1, 1, 1, 1]]


def test_round_half_even():
    import math
    ld = torch.tensor([math.log(1.5), math.log(2.5), math.log(3.5), -5.0, 0.0])
    # exp-1 = .5, 1.5, 2.5 -> half-to-even 0, 2, 2 ; negatives clamp to 0
    assert tte_oracle.round_durations(ld).tolist() == [0, 2, 2, 0, 0]


@pytest.mark.parametrize("name", ["voc_b1_t24", "voc_b2_t40", "voc_cfg1_t512", "voc_geom_x16_t300"])
def test_voc_oracle_matches_golden(golden_dir, name):
    g = _load(golden_dir, name)
    h = geometries.other_vocoder_config() if "geom" in name else configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=int(g["seed"]), weight_norm=True)
    wav = voc_oracle.code_generator_forward(sd, h, torch.from_numpy(g["code"]), torch.from_numpy(g["spkr"]))
    assert wav.shape == g["wav"].shape
    np.testing.assert_allclose(wav.numpy(), g["wav"], rtol=0, atol=2e-5)
    # folded-weight flavour of the checkpoint (after remove_weight_norm()) gives the same audio
    sd2 = synth.synth_vocoder_state_dict(h, seed=int(g["seed"]), weight_norm=False)
    wav2 = voc_oracle.code_generator_forward(sd2, h, torch.from_numpy(g["code"]), torch.from_numpy(g["spkr"]))
    np.testing.assert_allclose(wav2.numpy(), g["wav"], rtol=0, atol=2e-5)


def test_generator_oracle_matches_golden(golden_dir):
    g = _load(golden_dir, "gen_b1_t12")
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=int(g["seed"]), weight_norm=True)
    wav = voc_oracle.generator_forward(sd, h, torch.from_numpy(g["x"]))
    np.testing.assert_allclose(wav.numpy(), g["wav"], rtol=0, atol=2e-5)


def test_int16_truncates_toward_zero():
    y = torch.tensor([[0.5 / 32768 * 3, -0.5 / 32768 * 3, 0.99997]])
    assert voc_oracle.to_int16(y).tolist() == [1, -1, 32767]


def test_mel_oracle_matches_golden(golden_dir):
    """The golden file holds the output of the reference's own mel_spectrogram (utils/vocoder/dataset.py:43-69) with the oracle's
    mel basis standing in for the absent librosa call."""
    from oracle import mel_oracle
    g = _load(golden_dir, "mel_b2_n16384")
    mel = mel_oracle.mel_spectrogram(torch.from_numpy(g["y"]), 1024, 80, 16000, 256, 1024, 0, 8000)
    assert mel.shape == g["mel"].shape == (2, 80, 64)
    np.testing.assert_allclose(mel.numpy(), g["mel"], rtol=0, atol=1e-5)
    assert float(g["mel"].min()) == pytest.approx(np.log(1e-5), abs=1e-6)       # the clamp is exercised


def test_mel_filterbank_product_and_oracle_restatements_agree():
    """Two independent restatements of librosa.filters.mel's published algorithm (vectorised in the product, loop form in the
    oracle): identical float32 matrices, unit-area Slaney triangles."""
    from oracle import mel_oracle
    from parrot_tts_b200.vocoder.mel import mel_filterbank
    for sr, n_fft, n_mels, fmin, fmax in ((16000, 1024, 80, 0, 8000), (22050, 1024, 80, 0, None), (16000, 512, 40, 50, 7600)):
        a, b = mel_filterbank(sr, n_fft, n_mels, fmin, fmax), mel_oracle.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        assert a.dtype == np.float32 and a.shape == (n_mels, 1 + n_fft // 2)
        np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9)
        assert (a >= 0).all() and (a.sum(1) > 0).all()
