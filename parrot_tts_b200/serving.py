"""Batching front door for the vocoder.

The reference synthesises one utterance per call - ``DataLoader(batch_size=1)`` (inference.py:31-37) and
``Pool.imap(inference, idxs)`` over single items (utils/vocoder/inference.py:146-171,253-256) - then converts to int16 on the
host and peak-normalises with librosa before writing the wav (:71-73,169).  Here the same per-utterance results come out of
length-bucketed ragged batches: every item is masked to its own length inside the kernels (``lens=``), so it equals its
batch-of-1 result; the x32768 / int16 cast and the peak normalisation run on the device; utterances longer than
``long_form_units`` are streamed in chunks with a halo (BASELINE configs[4]); under ``torch.distributed`` each rank takes its
longest-processing-time-first share and nothing is exchanged on the data path.

Host logic only - every sample is produced by the CUDA library (no CPU fallback: CPU tensors raise in the ops)."""
import torch

from . import ops, sharding
from .vocoder.inference import stream_generate


def plan_batches(lengths, max_batch=8, max_pad_ratio=1.25, long_form_units=None, indices=None):
    """-> (batches, long_form, empty): ``batches`` = lists of utterance indices (longest first, shortest item within
    ``max_pad_ratio`` of the longest, at most ``max_batch`` items); ``long_form`` = indices streamed on their own;
    ``empty`` = zero-length utterances (no launch).  Pure function of the lengths: every rank computes the same plan
    without communicating."""
    idx = list(range(len(lengths))) if indices is None else list(indices)
    long_form = [i for i in idx if long_form_units is not None and lengths[i] > long_form_units]
    skip = set(long_form)
    rest = [i for i in idx if i not in skip]
    empty = [i for i in rest if lengths[i] == 0]
    rest = [i for i in rest if lengths[i] > 0]
    return sharding.length_buckets(rest, lengths, max_batch, max_pad_ratio), long_form, empty


def rank_share(lengths, world_size, rank):
    """Indices of the utterances rank ``rank`` synthesises (LPT on the vocoder cost; same on every rank)."""
    costs = [sharding.utterance_cost(n, with_tte=False) for n in lengths]
    return sorted(sharding.lpt_assign(costs, world_size)[rank])


class BatchSynthesizer:
    """``synth = BatchSynthesizer(generator); audio = synth(codes, spkrs)`` - list of unit sequences in, list of per-utterance
    audio out (same order).  ``generator`` is a ``parrot_tts_b200.vocoder.CodeGenerator`` on a CUDA device."""

    def __init__(self, generator, max_batch=8, max_pad_ratio=1.25, long_form_units=8192, chunk_units=1024, halo_units=32):
        self.generator = generator
        self.max_batch = int(max_batch)
        self.max_pad_ratio = float(max_pad_ratio)
        self.long_form_units = long_form_units
        self.chunk_units = int(chunk_units)
        self.halo_units = int(halo_units)
        self.last_plan = None
        self.last_launches = 0        # kernels of the library launched by the last call (vocoder forwards + PCM casts)

    def _device(self):
        return next(self.generator.parameters()).device

    def __call__(self, codes, spkrs=None, pcm16=True, normalize=False, indices=None, to_host=True):
        """codes: list of 1-D int sequences (lists / tensors) of unit ids; spkrs: list of speaker ids (or None for a
        single-speaker generator).  Returns a list aligned with ``codes`` (``None`` for utterances outside ``indices``):
        int16 PCM (``pcm16``), float32 peak-normalised audio (``normalize``: librosa.util.normalize of the int16 audio, as
        the reference writes it), or the raw float32 waveform; numpy arrays when ``to_host`` else device tensors."""
        if normalize and not pcm16:
            raise ValueError("normalize=True normalises the int16 audio (as the reference does): it needs pcm16=True")
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("BatchSynthesizer needs a generator on a CUDA device (there is no CPU path)")
        hop = self.generator.hop
        codes = [c.reshape(-1) if torch.is_tensor(c) and c.is_cuda else torch.as_tensor(c, dtype=torch.long).reshape(-1) for c in codes]
        lengths = [int(c.numel()) for c in codes]
        on_device = bool(codes) and all(c.is_cuda for c in codes)      # unit ids handed over on the device (Parrot.infer_device)
        batches, long_form, empty = plan_batches(lengths, self.max_batch, self.max_pad_ratio, self.long_form_units, indices)
        self.last_plan = {"batches": batches, "long_form": long_form, "empty": empty,
                          "units": sum(lengths[i] for b in batches for i in b),
                          "padded_units": sum(len(b) * lengths[b[0]] for b in batches)}
        out = [None] * len(codes)
        self.last_launches = 0

        def finish(wav):                               # f32 [n] on the device -> requested format
            if not pcm16:
                return wav
            pcm = ops.wav_to_pcm16(wav.contiguous())
            self.last_launches += 1 if wav.numel() else 0
            return ops.peak_normalize(pcm) if normalize else pcm

        def spk(i):
            return 0 if spkrs is None else int(spkrs[i])

        for b in batches:
            Lmax = lengths[b[0]]
            if on_device:                              # no host round trip of the ids
                batch_codes = torch.zeros(len(b), Lmax, dtype=torch.long, device=dev)
                for r, i in enumerate(b):
                    batch_codes[r, :lengths[i]] = codes[i]
            else:
                host = torch.zeros(len(b), Lmax, dtype=torch.long).pin_memory()
                for r, i in enumerate(b):
                    host[r, :lengths[i]] = codes[i]
                batch_codes = host.to(dev, non_blocking=True)
            kw = {"code": batch_codes,
                  "lens": torch.tensor([lengths[i] for i in b], dtype=torch.int32).to(dev, non_blocking=True)}
            if spkrs is not None:
                kw["spkr"] = torch.tensor([[spk(i)] for i in b], dtype=torch.long).to(dev, non_blocking=True)
            if pcm16:                                  # int16 straight from conv_post's epilogue (no f32 waveform round trip)
                pcm = self.generator.forward_pcm16(**kw)
                self.last_launches += self.generator.last_launches
                for r, i in enumerate(b):
                    p = pcm[r, 0, :lengths[i] * hop]
                    out[i] = ops.peak_normalize(p.contiguous()) if normalize else p
                continue
            wav = self.generator(**kw)                 # [len(b), 1, hop * Lmax], zero beyond each item's end
            self.last_launches += self.generator.last_launches
            for r, i in enumerate(b):
                out[i] = finish(wav[r, 0, :lengths[i] * hop])
        for i in long_form:                            # streamed: chunks of one utterance, each with its halo
            code = codes[i].to(dev).unsqueeze(0)
            sp = torch.tensor([[spk(i)]], dtype=torch.long, device=dev) if spkrs is not None else None
            parts = [w[0] for _, w in stream_generate(self.generator, code, sp, self.chunk_units, self.halo_units, pcm16=pcm16)]
            whole = torch.cat(parts)
            out[i] = (ops.peak_normalize(whole) if normalize else whole) if pcm16 else whole
        for i in empty:
            out[i] = finish(torch.zeros(0, device=dev))
        if to_host:
            out = [None if o is None else (ops.to_host(o) if o.numel() else o.cpu().numpy()) for o in out]
        return out


class SpeechPipeline:
    """``pipe = SpeechPipeline(parrot, generator); audio = pipe(phones, speakers)`` - character-token sequences in, per-utterance
    audio out: the whole hot path (demo.ipynb cells 11 + 13) for a pool of utterances, as one call.

    The reference infers one utterance per call (``DataLoader(batch_size=1)``, inference.py:31-37) because a padded batch changes every
    shorter item's result (the "positional" row is indexed by the PADDED length, modules/fft.py:18).  Here both stages run
    length-bucketed RAGGED batches whose items each equal their batch-of-1 result (``Parrot.infer_ragged``, ``CodeGenerator(lens=)``),
    and the unit ids go from the TTE to the vocoder ON THE DEVICE.  Under ``torch.distributed`` pass ``indices=rank_share_tte(...)``: each
    rank synthesises its longest-processing-time-first share, nothing is exchanged on the data path.
    """

    def __init__(self, parrot, generator, tte_max_batch=16, tte_max_pad_ratio=1.25, **synth_kwargs):
        self.parrot = parrot
        self.tte_max_batch, self.tte_max_pad_ratio = int(tte_max_batch), float(tte_max_pad_ratio)
        self.synth = BatchSynthesizer(generator, **synth_kwargs)
        self.last_units = None
        self.last_launches = 0

    def units(self, phones, speakers=None, indices=None):
        """-> list of 1-D i64 device tensors of unit ids (``None`` outside ``indices``).  Utterances are bucketed by character count
        and run through ``Parrot.infer_ragged``: every item equals its own batch-of-1 result."""
        dev = self.synth._device()
        idx = list(range(len(phones))) if indices is None else list(indices)
        out = [None] * len(phones)
        pad = self.parrot.src_pad_idx
        seqs = {i: torch.as_tensor(phones[i], dtype=torch.long).reshape(-1) for i in idx}
        n_chars = [int(seqs[i].numel()) if i in seqs else 0 for i in range(len(phones))]
        launches = 0
        live = [i for i in idx if n_chars[i] > 0]
        for b in sharding.length_buckets(live, n_chars, self.tte_max_batch, self.tte_max_pad_ratio):
            Smax = n_chars[b[0]]
            if all(seqs[i].is_cuda for i in b):
                ph = torch.full((len(b), Smax), pad, dtype=torch.long, device=dev)
                for r, i in enumerate(b):
                    ph[r, :n_chars[i]] = seqs[i]
            else:
                host = torch.full((len(b), Smax), pad, dtype=torch.long).pin_memory()
                for r, i in enumerate(b):
                    host[r, :n_chars[i]] = seqs[i]
                ph = host.to(dev, non_blocking=True)
            lens = torch.tensor([n_chars[i] for i in b], dtype=torch.int32).to(dev, non_blocking=True)
            spk = torch.tensor([0 if speakers is None else int(speakers[i]) for i in b], dtype=torch.long).to(dev, non_blocking=True)
            ids, n_units = self.parrot.infer_ragged(ph, lens, spk)
            launches += self.parrot.last_launches
            n_host = n_units.tolist()                  # one small D2H per bucket: the vocoder's batch plan needs the unit counts
            for r, i in enumerate(b):
                out[i] = ids[r, :n_host[r]]
        for i in idx:
            if out[i] is None:
                out[i] = torch.zeros(0, dtype=torch.long, device=dev)
        self.last_launches = launches
        return out

    def __call__(self, phones, speakers=None, vocoder_speakers=None, pcm16=True, normalize=False, indices=None, to_host=True):
        units = self.units(phones, speakers, indices)
        self.last_units = units
        idx = list(range(len(phones))) if indices is None else list(indices)
        vs = speakers if vocoder_speakers is None else vocoder_speakers
        audio = self.synth([units[i] for i in idx], None if vs is None else [vs[i] for i in idx], pcm16=pcm16, normalize=normalize,
                           to_host=to_host)
        out = [None] * len(phones)
        for i, a in zip(idx, audio):
            out[i] = a
        return out


def rank_share_tte(unit_lengths, world_size, rank):
    """Indices of the utterances rank ``rank`` runs end to end (LPT on the TTE + vocoder cost; same on every rank)."""
    costs = [sharding.utterance_cost(n, with_tte=True) for n in unit_lengths]
    return sorted(sharding.lpt_assign(costs, world_size)[rank])
