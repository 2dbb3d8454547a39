"""The whole demo path on the device: text -> cleaners -> DFATokenizer -> Parrot.infer -> BatchSynthesizer -> int16 PCM
(reference demo.ipynb cells 9, 11, 13), against the same chain on the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_text_to_speech_end_to_end(tmp_path):
    """text -> cleaners -> DFATokenizer -> Parrot.infer (per utterance) -> BatchSynthesizer -> int16 PCM, against the same chain on
    the oracle: unit ids bit-exact, PCM within the 1e-3 waveform bar (32.768 int16 steps, +1 for the truncation)."""
    import json
    import pickle
    import torch
    from oracle import tte_oracle, voc_oracle
    from parrot_tts_b200 import configs, synth
    from parrot_tts_b200.serving import BatchSynthesizer
    from parrot_tts_b200.text import TextToSpeech
    from parrot_tts_b200.tte import DFATokenizer, Parrot
    from parrot_tts_b200.vocoder import CodeGenerator
    symbols = list("abcdefghijklmnopqrstuvwxyz ,.")
    with open(tmp_path / "symbols.pkl", "wb") as f:
        pickle.dump(symbols, f)
    with open(tmp_path / "speakers.json", "w") as f:
        json.dump({f"s{i}": i for i in range(10)}, f)
    tok = DFATokenizer(tmp_path)
    cfg = configs.default_tte_config(root_path=str(tmp_path))
    sd_t = synth.synth_tte_state_dict(cfg, len(tok), tok.pad_idx, 10, seed=4)
    tte = Parrot(cfg, len(tok), tok.pad_idx)
    tte.load_state_dict(sd_t)
    tte = tte.eval().to("cuda:0")
    h = configs.default_vocoder_config()
    sd_v = synth.synth_vocoder_state_dict(h, seed=4)
    gen = CodeGenerator(h)
    gen.load_state_dict(sd_v)
    gen = gen.to("cuda:0")
    tts = TextToSpeech(tok, symbols, tte.infer, BatchSynthesizer(gen), device="cuda:0")
    texts, speakers = ["Dr. Who has 2 cats.", "???", "hello there, general"], [1, 2, 3]
    audio = tts(texts, speakers)
    assert audio[1] is None and tts.last_units[1] is None
    for i in (0, 2):
        phones = torch.tensor([tok.tokenize(tts.last_characters[i])])
        ids_ref, _ = tte_oracle.parrot_infer(sd_t, {"phones": phones, "src_mask": phones != 0, "speaker": torch.tensor([speakers[i]])})
        assert tts.last_units[i] == ids_ref[0]
        ref = voc_oracle.code_generator_forward(sd_v, h, torch.tensor([ids_ref[0]]), torch.tensor([[speakers[i]]]))
        pcm_ref = voc_oracle.to_int16(ref)
        assert audio[i].dtype == np.int16 and audio[i].shape == pcm_ref.shape
        assert int(np.abs(audio[i].astype(np.int32) - pcm_ref.astype(np.int32)).max()) <= 34
