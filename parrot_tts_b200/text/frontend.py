"""Text front end of the demo path: text -> cleaned text -> aligner characters -> ``batch`` for ``Parrot.infer`` -> unit
ids -> ``CodeGenerator`` (reference demo.ipynb cells 9, 11, 13).  Host string work plus two calls into the device path;
no device math lives here.
"""
import unicodedata

import torch

from . import cleaners


def looks_english(text):
    """Stand-in for ``langdetect.detect(text) == 'en'`` (demo.ipynb cell 9; ``langdetect`` is not installed): the demo only
    uses the answer to pick the English cleaner, so the question asked here is the one that matters for that choice -
    is every letter of the text Latin script?"""
    letters = [c for c in text if unicodedata.category(c)[0] == "L"]
    return bool(letters) and all(unicodedata.name(c, "").startswith("LATIN") for c in letters)


def pick_cleaner(text, transliteration=False, english=None):
    """demo.ipynb cell 9: English text -> english_cleaners; otherwise by the ``transliteration`` switch."""
    if english is None:
        english = looks_english(text)
    if english:
        return cleaners.english_cleaners
    return cleaners.nonenglish_cleaners if transliteration else cleaners.nonenglish_cleaners_no_transliteration


def text_to_characters(text, symbols, transliteration=False, english=None):
    """Cleaned text -> the aligner's character symbols: characters the symbol table does not know are dropped, the blank
    becomes ``sil`` (demo.ipynb cell 9).  ``symbols`` is anything supporting ``in`` (the ``symbols.pkl`` table)."""
    cleaned = pick_cleaner(text, transliteration, english)(text)
    return ["sil" if c == " " else c for c in cleaned if c in symbols]


def characters_to_batch(characters, tokenizer, speaker=0):
    """One utterance -> the ``batch`` dict ``Parrot.infer`` takes (demo.ipynb cell 11; keys of reference
    modules/data.py:102-120 that inference reads: phones, src_mask, speaker)."""
    phones = torch.tensor([tokenizer.tokenize(characters)], dtype=torch.long)
    return {"ids": ["text"], "phones": phones, "src_mask": phones != tokenizer.pad_idx,
            "speaker": torch.tensor([int(speaker)], dtype=torch.long)}


class TextToSpeech:
    """``tts = TextToSpeech(tokenizer, symbols, units_fn, audio_fn); audio = tts(["some text", ...], speakers=[0, ...])``

    units_fn(batch) -> list[list[int]]     e.g. ``LitParrot.infer`` - called once per utterance, batch of one, which is the
                                           reference's own inference mode (inference.py:34) and the only batching whose
                                           result does not depend on the other items
    audio_fn(codes, spkrs) -> list         e.g. a ``serving.BatchSynthesizer`` - called once for all utterances, so the
                                           vocoder runs length-bucketed batches
    Utterances whose text cleans down to nothing yield ``None``.
    """

    def __init__(self, tokenizer, symbols, units_fn, audio_fn, transliteration=False, device=None):
        self.tokenizer, self.symbols = tokenizer, symbols
        self.units_fn, self.audio_fn = units_fn, audio_fn
        self.transliteration = transliteration
        self.device = device
        self.last_characters = None
        self.last_units = None

    def __call__(self, texts, speakers=None, vocoder_speakers=None):
        """speakers: TTE speaker id per text (default 0); vocoder_speakers: vocoder voice per text (default = speakers)."""
        texts = [texts] if isinstance(texts, str) else list(texts)
        speakers = [0] * len(texts) if speakers is None else [int(s) for s in speakers]
        vocoder_speakers = speakers if vocoder_speakers is None else [int(s) for s in vocoder_speakers]
        if not len(texts) == len(speakers) == len(vocoder_speakers):
            raise ValueError("texts, speakers and vocoder_speakers must have the same length")
        chars = [text_to_characters(t, self.symbols, self.transliteration) for t in texts]
        units = []
        for c, s in zip(chars, speakers):
            if not c:
                units.append(None)
                continue
            batch = characters_to_batch(c, self.tokenizer, s)
            if self.device is not None:
                batch = {k: v.to(self.device) if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
            units.append(list(self.units_fn(batch)[0]))
        self.last_characters, self.last_units = chars, units
        live = [i for i, u in enumerate(units) if u]
        audio = self.audio_fn([units[i] for i in live], [vocoder_speakers[i] for i in live]) if live else []
        out = [None] * len(texts)
        for i, a in zip(live, audio):
            out[i] = a
        return out
