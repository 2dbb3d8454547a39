"""TTE model with the reference's constructor / forward / infer signatures and state_dict layout
(reference modules/parrot.py:12-120, inference.py:10-23), executing on libparrot_b200.so."""
import ctypes as C
import json
import os

import torch
import torch.nn as nn

from .. import _lib
from ._engine import TteEngine, need_cuda
from .._weights import WeightStamp
from .duration import DurationPredictor
from .fft import FFTBlock, SinusoidalPosEmb


class Parrot(nn.Module):
    #: pb_tte_cfg.gemm_mode: 1 = tcgen05 (projections, FFN convolutions, head and the fused attention on split-fp16 operands with
    #: fp32 accumulation - fp32-class accuracy, bit-exact ids / durations on every golden vector; the default),
    #: 0 = fp32 SIMT kernels (also used for the unfolded-projection reference mode).
    gemm_mode = 1
    #: fold mha.in_proj . qkv and wo . out_proj at load (exact in real arithmetic; 6 GEMMs -> 2 per block)
    fold_projections = True

    def __init__(self, data_config, src_vocab_size, src_pad_idx):
        super().__init__()
        tr = data_config["transformer"]
        dp = data_config["duration_predictor"]
        self.max_len = tr["max_len"]
        self.d_model = tr["d_model"]
        self.src_vocab_size = src_vocab_size
        self.src_pad_idx = src_pad_idx
        self.n_codes = data_config["preprocess"]["hubert_codes"]
        self._tr = tr
        self._dp = dp

        self.pos_emb = SinusoidalPosEmb(self.max_len, self.d_model)
        self.tok_emb = nn.Embedding(src_vocab_size, self.d_model, src_pad_idx)
        self.speaker_emb = None
        with open(os.path.join(data_config["path"]["root_path"], "speakers.json")) as f:
            n_speaker = len(json.load(f))
        if n_speaker > 1:
            self.speaker_emb = nn.Embedding(n_speaker, self.d_model)
        self.duration_predictor = DurationPredictor(self.d_model, dp["n_filter"], dp["kernel_size"], dp["dropout_p"])

        def stack(side):
            return nn.ModuleList(
                FFTBlock(self.d_model, tr[side]["n_head"], tr["conv_n_filter"], tr["conv_kernel_sizes"], tr[side]["dropout_p"])
                for _ in range(tr[side]["n_layer"]))
        self.encoder_layers = stack("encoder")
        self.decoder_layers = stack("decoder")
        self.head = nn.Linear(self.d_model, self.n_codes)
        self._eng = TteEngine()
        self._wstamp = WeightStamp(self)
        self.last_launches = 0

    # ------------------------------------------------------------------ engine
    def _cfg_fields(self):
        tr, dp = self._tr, self._dp
        if tr["encoder"]["n_head"] != tr["decoder"]["n_head"]:
            raise NotImplementedError("encoder and decoder must use the same number of heads")
        return dict(src_vocab_size=self.src_vocab_size, src_pad_idx=self.src_pad_idx,
                   n_speaker=self.speaker_emb.num_embeddings if self.speaker_emb is not None else 0,
                   d_model=self.d_model, n_head=tr["encoder"]["n_head"], n_enc_layer=len(self.encoder_layers),
                   n_dec_layer=len(self.decoder_layers), conv_n_filter=tr["conv_n_filter"], conv_k1=tr["conv_kernel_sizes"][0],
                   conv_k2=tr["conv_kernel_sizes"][1], dp_n_filter=dp["n_filter"], dp_kernel=dp["kernel_size"],
                   n_codes=self.n_codes, max_len=self.max_len, fold_projections=int(self.fold_projections),
                   gemm_mode=int(self.gemm_mode) if self.fold_projections else 0)   # the unfolded reference mode is fp32 SIMT

    def _handle(self, device):
        return self._eng.get(self._cfg_fields(), self._wstamp, device)

    def invalidate_weights(self):
        """Force the next call to re-upload, re-fold and re-pack the weights.  Needed only after edits torch cannot see (``p.data.copy_()``
        and the like bump no version counter); ``load_state_dict``, in-place ops on the parameters and ``.to()`` are detected."""
        self._wstamp.invalidate()

    # ---- packed-arena hand-over (sharding.broadcast_engines), as on the vocoder
    def engine_arena(self, device, receive=False):
        lib = _lib.load()
        device = torch.device(device)
        hnd = self._eng.create_empty(self._cfg_fields(), device) if receive else self._handle(device)
        self._arena_device = device
        n = C.c_int64(0)
        _lib.check(lib.pb_tte_arena_bytes(hnd, C.byref(n)))
        return int(lib.pb_tte_arena_ptr(hnd)), int(n.value)

    def engine_adopt(self):
        self._eng.adopt(self._cfg_fields(), self._wstamp, self._arena_device)

    def _encode(self, batch, dur_in=None, src_lens=None):
        phones = batch["phones"]
        need_cuda(phones, "Parrot.forward")
        if self.training:
            raise _lib.ParrotB200Error("Parrot is inference-only on this backend: call .eval() first (dropout is not implemented)")
        dev = phones.device
        hnd = self._handle(dev)
        B, S = phones.shape
        phones = phones.to(torch.int64).contiguous()
        src_mask = batch["src_mask"].to(device=dev, dtype=torch.uint8).contiguous()
        speaker = None
        if self.speaker_emb is not None:
            speaker = batch["speaker"].to(device=dev, dtype=torch.int64).reshape(-1).contiguous()
        log_dur = torch.empty(B, S, device=dev, dtype=torch.float32)
        dur = torch.empty(B, S, device=dev, dtype=torch.int64)
        out_lens = torch.empty(B, device=dev, dtype=torch.int64)
        if dur_in is not None:
            dur_in = dur_in.to(device=dev, dtype=torch.int64).contiguous()
        if src_lens is not None:
            src_lens = src_lens.to(device=dev, dtype=torch.int32).contiguous()
        T = C.c_int32(0)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pb_tte_encode(hnd, _lib.ptr(phones), _lib.ptr(src_mask), _lib.ptr(speaker), _lib.ptr(dur_in),
                                                 _lib.ptr(src_lens), B, S, _lib.ptr(log_dur), _lib.ptr(dur), _lib.ptr(out_lens),
                                                 C.byref(T), _lib.stream_ptr(dev)))
        self.last_launches = int(_lib.load().pb_tte_last_launches(hnd))
        return hnd, dev, B, S, int(T.value), log_dur, dur, out_lens

    def _decode(self, hnd, dev, B, S, T, want_logits, want_ids, tgt_mask_in=None):
        logits = torch.empty(B, T, self.n_codes, device=dev, dtype=torch.float32) if want_logits else None
        ids = torch.empty(B, T, device=dev, dtype=torch.int64) if want_ids else None
        tgt_mask = torch.empty(B, T, device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().pb_tte_decode(hnd, B, S, T, _lib.ptr(tgt_mask_in), _lib.ptr(logits), _lib.ptr(ids), _lib.ptr(tgt_mask),
                                                 _lib.stream_ptr(dev)))
        self.last_launches += int(_lib.load().pb_tte_last_launches(hnd))
        return logits, ids, tgt_mask

    # ------------------------------------------------------------------ reference API
    def forward(self, batch, inference=False):
        """-> (logits f32 [B,T,n_codes], src_mask, tgt_mask bool [B,T], log_dur f32 [B,S])  (reference parrot.py:90-110).
        inference=False uses the teacher durations ``batch['duration']`` for the length regulator."""
        dur_in = None
        if inference is not True:
            assert "duration" in batch.keys()
            dur_in = batch["duration"]
        hnd, dev, B, S, T, log_dur, _, _ = self._encode(batch, dur_in)
        given = batch.get("tgt_mask") if dur_in is not None else None
        mask_in = None
        if given is not None:
            # the reference's length_regulator returns a supplied mask unchanged and the decoder attends with it (duration.py:9-10,22-24)
            assert given.shape[1] == T
            mask_in = given.to(device=dev, dtype=torch.uint8).contiguous()
        logits, _, tgt_mask = self._decode(hnd, dev, B, S, T, True, False, mask_in)
        return logits, batch["src_mask"], (given if given is not None else tgt_mask.bool()), log_dur

    def infer_device(self, batch):
        """Device-resident variant of infer(): (ids i64 [B,T], n_valid i64 [B]) where item b's units are
        ids[b, :n_valid[b]] - nothing is copied to the host."""
        hnd, dev, B, S, T, _, _, out_lens = self._encode(batch)
        _, ids, _ = self._decode(hnd, dev, B, S, T, False, True)
        return ids, torch.clamp(out_lens + 1, max=T)

    def infer_ragged(self, phones, lens, speaker=None, return_logits=False):
        """Batched inference in which every item equals the reference's result for that item ALONE (its batch-of-1 inference mode,
        inference.py:31-37): phones i64 [B, S] padded with ``src_pad_idx``, lens [B] = characters per item, speaker i64 [B].
        -> (ids i64 [B, T], n_units i64 [B]), item b's units are ``ids[b, :n_units[b]]`` (exactly its own unit count - the extra
        `<=` position of a padded batch does not exist for an item run alone).

        In the reference a padded batch changes every shorter item's result (the "positional" row is indexed by the PADDED length,
        modules/fft.py:18; padded positions feed the k = 9 convolutions; the target mask is ``t <= len``, modules/data.py:18), which
        is why it infers one utterance at a time.  Here the kernels take each item's own length: its own ``pe`` rows, zeros past its
        own end wherever a convolution looks, keys past its end masked - so a length-bucketed batch reproduces the one-at-a-time
        results at batch throughput.  ``return_logits=True`` appends the logits f32 [B, T, n_codes] (rows past an item's end are
        meaningless)."""
        need_cuda(phones, "Parrot.infer_ragged")
        dev = phones.device
        lens = torch.as_tensor(lens).to(device=dev, dtype=torch.int32).reshape(-1)
        B, S = phones.shape
        if lens.numel() != B:
            raise ValueError(f"lens must hold one length per item, got {lens.numel()} for batch {B}")
        src_mask = torch.arange(S, device=dev).unsqueeze(0) < lens.unsqueeze(1)
        batch = {"phones": phones, "src_mask": src_mask, "speaker": speaker if speaker is not None else torch.zeros(B, dtype=torch.long, device=dev)}
        hnd, dev, B, S, T, _, _, out_lens = self._encode(batch, src_lens=lens)
        if T == 0:
            return torch.zeros(B, 0, dtype=torch.int64, device=dev), out_lens
        logits, ids, _ = self._decode(hnd, dev, B, S, T, bool(return_logits), True)
        return (ids, out_lens, logits) if return_logits else (ids, out_lens)

    def infer(self, batch):
        """-> list[list[int]]: per item the argmax unit ids at the positions its tgt_mask marks (reference parrot.py:112-120)."""
        assert self.training is False
        ids, n_valid = self.infer_device(batch)
        ids_h, n_h = ids.cpu(), n_valid.cpu().tolist()
        return [ids_h[b, :n].tolist() for b, n in enumerate(n_h)]


class LitParrot(nn.Module):
    """Stand-in for the reference's LightningModule wrapper (inference.py:10-23): same attribute (`parrot`), same
    checkpoint layout (`hyper_parameters` + `state_dict` with the ``parrot.`` prefix), no Lightning dependency."""

    def __init__(self, data_config, src_vocab_size, src_pad_idx):
        super().__init__()
        self.hparams = dict(data_config=data_config, src_vocab_size=src_vocab_size, src_pad_idx=src_pad_idx)
        self.parrot = Parrot(data_config, src_vocab_size, src_pad_idx)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, weights_only=True, strict=True, **overrides):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=weights_only)
        hp = dict(ckpt["hyper_parameters"])
        hp.update(overrides)
        model = cls(hp["data_config"], hp["src_vocab_size"], hp["src_pad_idx"])
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        return model.eval()

    def save_checkpoint(self, path):
        torch.save({"hyper_parameters": self.hparams, "state_dict": self.state_dict()}, path)

    def infer(self, batch):
        self.eval()
        return self.parrot.infer(batch)
