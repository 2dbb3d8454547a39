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
    shorter item's result (the "positional" row is indexed by the PADDED length, modules/fft.py:18).  Here a BUCKET of utterances of
    similar length goes through both stages as one ragged batch whose items each equal their batch-of-1 result
    (``Parrot.infer_ragged`` -> ``CodeGenerator(lens=)``): the padded unit ids and the per-item unit counts go from the TTE to the
    vocoder ON THE DEVICE as they are (no re-bucketing, no host round trip), int16 PCM comes out of ``conv_post``'s epilogue, and the
    host reads the unit counts only after every bucket has been launched.  Under ``torch.distributed`` pass
    ``batches=plan_shares(...)[rank]``: each rank synthesises its share of the buckets, nothing is exchanged on the data path.
    """

    def __init__(self, parrot, generator, max_batch=8, max_pad_ratio=1.1, **synth_kwargs):
        self.parrot = parrot
        self.generator = generator
        self.max_batch, self.max_pad_ratio = int(max_batch), float(max_pad_ratio)
        self.synth = BatchSynthesizer(generator, max_batch=max_batch, **synth_kwargs)      # the vocoder-only front door (kept for callers)
        self.last_units = None
        self.last_launches = 0
        self.last_batches = None
        self.on_vocoder_forward = None      # optional callback(bucket indices, padded units) after each vocoder launch (profiling hooks)
        self._copy_stream = None

    def _device(self):
        return next(self.generator.parameters()).device

    def plan(self, phones, indices=None):
        """Buckets (lists of utterance indices, longest first, shortest item within ``max_pad_ratio`` of the longest, at most
        ``max_batch`` items) over the utterances in ``indices`` - by character count, the only length known before the TTE ran."""
        idx = list(range(len(phones))) if indices is None else list(indices)
        n_chars = [int(torch.as_tensor(p).numel()) for p in phones]
        return sharding.length_buckets([i for i in idx if n_chars[i] > 0], n_chars, self.max_batch, self.max_pad_ratio)

    def __call__(self, phones, speakers=None, vocoder_speakers=None, pcm16=True, normalize=False, indices=None, to_host=True, batches=None):
        """phones: list of 1-D char-token sequences (host or device); speakers / vocoder_speakers: TTE speaker and vocoder voice per
        utterance.  ``batches`` (explicit buckets, e.g. a rank's share from ``plan_shares``) or ``indices`` select the utterances; the
        result list is aligned with ``phones`` (``None`` for the others).  Audio per utterance: int16 PCM (``pcm16``), float32
        peak-normalised (``normalize``) or the raw float32 waveform; numpy arrays when ``to_host`` else device tensors."""
        if normalize and not pcm16:
            raise ValueError("normalize=True normalises the int16 audio (as the reference does): it needs pcm16=True")
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("SpeechPipeline needs its models on a CUDA device (there is no CPU path)")
        explicit = batches is not None
        if not explicit:
            batches = self.plan(phones, indices)
        vs = speakers if vocoder_speakers is None else vocoder_speakers
        pad, hop = self.parrot.src_pad_idx, self.generator.hop
        multispkr = bool(getattr(self.generator, "multispkr", None))
        launches, staged, pending = 0, [], []
        # phase 1: the TTE of every bucket (its encode reads the unit count back - the only synchronisation on the path - while the
        # queue holds nothing but short TTE work); phase 2: the vocoder of every bucket, back to back with no synchronisation at all
        for b in batches:
            seqs = [torch.as_tensor(phones[i], dtype=torch.long).reshape(-1) for i in b]
            n_chars = [int(q.numel()) for q in seqs]
            Smax = max(n_chars)
            if all(q.is_cuda for q in seqs):
                ph = torch.full((len(b), Smax), pad, dtype=torch.long, device=dev)
                for r, q in enumerate(seqs):
                    ph[r, :n_chars[r]] = q
            else:
                host = torch.full((len(b), Smax), pad, dtype=torch.long).pin_memory()
                for r, q in enumerate(seqs):
                    host[r, :n_chars[r]] = q
                ph = host.to(dev, non_blocking=True)
            lens = torch.tensor(n_chars, dtype=torch.int32).to(dev, non_blocking=True)
            spk = torch.tensor([0 if speakers is None else int(speakers[i]) for i in b], dtype=torch.long).to(dev, non_blocking=True)
            ids, n_units = self.parrot.infer_ragged(ph, lens, spk)            # ids i64 [B, Tmax] padded, n_units i64 [B]: stay on the device
            launches += self.parrot.last_launches
            staged.append((b, ids, n_units))
        for b, ids, n_units in staged:
            out = None
            if ids.shape[1] > 0:
                kw = {"code": ids, "lens": n_units.to(torch.int32)}
                if multispkr:
                    kw["spkr"] = torch.tensor([[0 if vs is None else int(vs[i])] for i in b], dtype=torch.long).to(dev, non_blocking=True)
                out = self.generator.forward_pcm16(**kw) if pcm16 else self.generator(**kw)      # [B, 1, hop * Tmax], zero past each item's end
                launches += self.generator.last_launches
                if self.on_vocoder_forward is not None:
                    self.on_vocoder_forward(b, int(ids.shape[0] * ids.shape[1]))
            # the bucket's audio block and unit counts leave for pinned host memory on a copy stream, behind the next buckets' kernels
            host_block = host_n = None
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(dev)
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(done)
                host_n = torch.empty(n_units.shape, dtype=n_units.dtype, pin_memory=True)
                host_n.copy_(n_units, non_blocking=True)
                n_units.record_stream(self._copy_stream)
                if to_host and out is not None and not normalize:
                    host_block = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                    host_block.copy_(out, non_blocking=True)
                    out.record_stream(self._copy_stream)
            pending.append((b, ids, host_n, out, host_block))
        # everything is launched: only now wait for the unit counts and cut the utterances out
        if self._copy_stream is not None:
            self._copy_stream.synchronize()
        result = [None] * len(phones)
        units = [None] * len(phones)
        for b, ids, host_n, out, host_block in pending:
            n_host = host_n.tolist()
            host_block = host_block.numpy() if host_block is not None else None
            for r, i in enumerate(b):
                n = int(n_host[r])
                units[i] = ids[r, :n]
                if out is None or n == 0:
                    a = torch.zeros(0, dtype=torch.int16 if pcm16 and not normalize else torch.float32)
                    result[i] = a.numpy() if to_host else a.to(dev)
                elif host_block is not None:
                    result[i] = host_block[r, 0, :n * hop]
                else:
                    a = out[r, 0, :n * hop]
                    if normalize:
                        a = ops.peak_normalize(a.contiguous())
                        launches += 2
                    result[i] = ops.to_host(a) if to_host else a
        if not explicit:       # utterances of the selection that have no characters: empty audio, no launch
            wanted = range(len(phones)) if indices is None else indices
            covered = {i for b in batches for i in b}
            for i in wanted:
                if i not in covered and torch.as_tensor(phones[i]).numel() == 0:
                    units[i] = torch.zeros(0, dtype=torch.long, device=dev)
                    a = torch.zeros(0, dtype=torch.int16 if pcm16 and not normalize else torch.float32)
                    result[i] = a.numpy() if to_host else a.to(dev)
        self.last_units, self.last_launches, self.last_batches = units, launches, [list(b) for b in batches]
        return result

    def units(self, phones, speakers=None, indices=None):
        """TTE only: -> list of 1-D i64 device tensors of unit ids (``None`` outside ``indices``)."""
        dev = self._device()
        pad = self.parrot.src_pad_idx
        out = [None] * len(phones)
        launches = 0
        for b in self.plan(phones, indices):
            seqs = [torch.as_tensor(phones[i], dtype=torch.long).reshape(-1) for i in b]
            n_chars = [int(q.numel()) for q in seqs]
            ph = torch.full((len(b), max(n_chars)), pad, dtype=torch.long, device=dev)
            for r, q in enumerate(seqs):
                ph[r, :n_chars[r]] = q.to(dev)
            spk = torch.tensor([0 if speakers is None else int(speakers[i]) for i in b], dtype=torch.long, device=dev)
            ids, n_units = self.parrot.infer_ragged(ph, torch.tensor(n_chars, dtype=torch.int32, device=dev), spk)
            launches += self.parrot.last_launches
            for r, (i, n) in enumerate(zip(b, n_units.tolist())):
                out[i] = ids[r, :int(n)]
        self.last_launches = launches
        return out


def rank_share_tte(unit_lengths, world_size, rank):
    """Indices of the utterances rank ``rank`` runs end to end (LPT on the TTE + vocoder cost; same on every rank)."""
    costs = [sharding.utterance_cost(n, with_tte=True) for n in unit_lengths]
    return sorted(sharding.lpt_assign(costs, world_size)[rank])


def plan_shares(unit_estimates, world_size, max_batch=8, max_pad_ratio=1.1):
    """-> per rank, a list of buckets (lists of utterance indices): the pool is bucketed GLOBALLY by length first (so that every
    bucket is full and tightly padded whatever the number of ranks), then the buckets go to the ranks longest-processing-time-first
    on their padded cost.  Splitting utterance by utterance instead leaves every rank a thin mix of all lengths: at 8 ranks the
    256-utterance pool of BASELINE configs[3] then pays 9.6 % of padded work on the slowest rank against 2.6 % this way.
    ``unit_estimates``: expected unit count per utterance (e.g. 8 x its characters); a pure function, identical on every rank."""
    idx = [i for i in range(len(unit_estimates)) if unit_estimates[i] > 0]
    buckets = sharding.length_buckets(idx, unit_estimates, max_batch, max_pad_ratio)
    costs = [len(b) * sharding.utterance_cost(max(unit_estimates[i] for i in b), with_tte=True) for b in buckets]
    assign = sharding.lpt_assign(costs, world_size)
    return [[buckets[k] for k in sorted(assign[r])] for r in range(world_size)]
