"""ORACLE (test infrastructure, not product code): CPU restatement of the reference TTE path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The product (``parrot_tts_b200``) never does.

What it restates (all citations relative to the reference tree):
  * Parrot.forward / Parrot.infer            modules/parrot.py:90-120
  * SinusoidalPosEmb.forward                 modules/fft.py:17-19   (adds the single row pe[L])
  * Attention.forward + nn.MultiheadAttention modules/fft.py:53-59; torch/nn/functional.py
                                             multi_head_attention_forward (q/k/v differ => three
                                             separate in_proj thirds, scale on q, additive -inf
                                             key-padding mask, softmax, bmm, out_proj)
  * ConvLayer / FFTBlock                     modules/fft.py:78-82, 94-100
  * DurationPredictor / Conv                 modules/duration.py:41-48, 74-79
  * duration rounding                        modules/parrot.py:82-86 (round-half-even, clamp >= 0)
  * length_regulator                         modules/duration.py:6-24
  * get_mask_from_lengths                    modules/data.py:8-20  (``ids <= len``: len+1 valid)

Arithmetic is torch fp32 on CPU (the library the reference itself calls), written functionally
over a flat state_dict so it needs neither the reference classes nor ``speakers.json``.

Parity pin: this restatement is checked against the *reference's own modules* executed in the
build container (tests/test_oracle_vs_reference.py, skipped where /root/reference is absent)
and against the committed golden vectors produced by oracle/gen_golden.py from those modules
(tests/golden/*.npz).  The reference itself ships no tests or golden vectors.
"""
import math

import torch
import torch.nn.functional as F


def get_mask_from_lengths(lengths, max_len):
    # modules/data.py:8-20 - note `<=` (off by one, part of the parity contract)
    lengths = torch.as_tensor(lengths, dtype=torch.long)
    ids = torch.arange(0, int(max_len)).unsqueeze(0)
    return ids <= lengths.unsqueeze(1)


def pos_emb(sd, x):
    # modules/fft.py:17-19 - raises IndexError when x.size(1) >= max_len
    return sd["pos_emb.pe"][x.size(1)] + x


def attention(sd, p, x, key_padding_mask, n_head):
    """modules/fft.py:53-59 -> F.multi_head_attention_forward (unpacked q/k/v path)."""
    B, L, D = x.shape
    dh = D // n_head
    qkv = x @ sd[p + ".qkv.weight"].t()
    q, k, v = qkv.split(D, dim=2)
    w_q, w_k, w_v = sd[p + ".mha.in_proj_weight"].chunk(3)
    q = q @ w_q.t()
    k = k @ w_k.t()
    v = v @ w_v.t()
    # [B, L, D] -> [B*H, L, dh]  (batch_first handled by transposes inside nn.MultiheadAttention)
    def heads(t):
        return t.transpose(0, 1).reshape(L, B * n_head, dh).transpose(0, 1)
    q, k, v = heads(q), heads(k), heads(v)
    q = q * math.sqrt(1.0 / float(dh))
    if key_padding_mask is not None:
        m = torch.zeros(B, L, dtype=x.dtype, device=x.device).masked_fill_(key_padding_mask, float("-inf"))
        m = m.view(B, 1, 1, L).expand(-1, n_head, -1, -1).reshape(B * n_head, 1, L)
        s = torch.baddbmm(m, q, k.transpose(-2, -1))
    else:
        s = torch.bmm(q, k.transpose(-2, -1))
    s = F.softmax(s, dim=-1)
    o = torch.bmm(s, v)  # [B*H, L, dh]
    o = o.transpose(0, 1).reshape(L, B, D).transpose(0, 1)
    o = o @ sd[p + ".mha.out_proj.weight"].t()
    return o @ sd[p + ".wo.weight"].t()


def conv_layer(sd, p, x):
    # modules/fft.py:78-82
    k1 = sd[p + ".conv1.weight"].shape[2]
    k2 = sd[p + ".conv2.weight"].shape[2]
    o = x.transpose(1, 2)
    o = F.conv1d(o, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=(k1 - 1) // 2)
    o = F.conv1d(F.relu(o), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=(k2 - 1) // 2)
    return o.transpose(1, 2)


def fft_block(sd, p, x, key_padding_mask, n_head):
    # modules/fft.py:94-100 (pre-LN, eps 1e-5)
    D = x.shape[-1]
    xn = F.layer_norm(x, (D,), sd[p + ".attn_norm.weight"], sd[p + ".attn_norm.bias"], 1e-5)
    h = x + attention(sd, p + ".attention", xn, key_padding_mask, n_head)
    hn = F.layer_norm(h, (D,), sd[p + ".conv_norm.weight"], sd[p + ".conv_norm.bias"], 1e-5)
    return h + conv_layer(sd, p + ".convlayer", hn)


def duration_predictor(sd, x, mask):
    # modules/duration.py:41-48; second conv's padding is hard-coded 1 (:34)
    p = "duration_predictor"
    k = sd[p + ".layers.0.conv.weight"].shape[2]
    nf = sd[p + ".layers.0.conv.weight"].shape[0]
    o = F.conv1d(x.transpose(1, 2), sd[p + ".layers.0.conv.weight"], sd[p + ".layers.0.conv.bias"],
                 padding=(k - 1) // 2).transpose(1, 2)
    o = F.layer_norm(F.relu(o), (nf,), sd[p + ".layers.2.weight"], sd[p + ".layers.2.bias"], 1e-5)
    o = F.conv1d(o.transpose(1, 2), sd[p + ".layers.4.conv.weight"], sd[p + ".layers.4.conv.bias"],
                 padding=1).transpose(1, 2)
    o = F.layer_norm(F.relu(o), (nf,), sd[p + ".layers.6.weight"], sd[p + ".layers.6.bias"], 1e-5)
    o = (o @ sd[p + ".proj.weight"].t() + sd[p + ".proj.bias"]).squeeze(-1)
    if mask is not None:
        o = o.masked_fill(mask, 0.0)
    return o


def round_durations(log_dur):
    # modules/parrot.py:82-86
    return torch.clamp(torch.round(torch.exp(log_dur) - 1), min=0).long()


def length_regulator(batch_seq, batch_dur):
    # modules/duration.py:6-24
    max_len = int(batch_dur.sum(dim=1).max())
    expanded, out_lens = [], []
    for seq, dur in zip(batch_seq, batch_dur):
        seq = seq.repeat_interleave(dur, dim=0)
        out_lens.append(seq.shape[0])
        expanded.append(F.pad(seq, (0, 0, 0, max_len - seq.shape[0]), "constant", 0.0))
    return torch.stack(expanded), get_mask_from_lengths(out_lens, max_len).to(batch_seq.device), out_lens


def _n_layers(sd, stack):
    n = 0
    while f"{stack}.{n}.attn_norm.weight" in sd:
        n += 1
    return n


def parrot_forward(sd, batch, n_head=2, dur_override=None, tgt_mask=None):
    """Parrot.forward(batch, inference=True) (modules/parrot.py:90-110).

    Returns dict(logits, src_mask, tgt_mask, log_dur, dur, out_lens, enc_out).
    ``dur_override`` plays the role of ``batch['duration']`` (teacher-forced path, :102) and ``tgt_mask`` of ``batch['tgt_mask']``:
    length_regulator returns a supplied mask unchanged (modules/duration.py:9-10,22-24) and the decoder attends with it (:107).
    """
    with torch.no_grad():
        src_mask = batch["src_mask"]
        out = F.embedding(batch["phones"], sd["tok_emb.weight"])
        out = pos_emb(sd, out)
        for i in range(_n_layers(sd, "encoder_layers")):
            out = fft_block(sd, f"encoder_layers.{i}", out, ~src_mask, n_head)
        if "speaker_emb.weight" in sd:
            out = out + F.embedding(batch["speaker"], sd["speaker_emb.weight"]).unsqueeze(1)
        enc_out = out
        log_dur = duration_predictor(sd, out, ~src_mask)
        dur = round_durations(log_dur) if dur_override is None else dur_override
        given = tgt_mask
        out, tgt_mask, out_lens = length_regulator(out, dur)
        if given is not None:
            assert given.shape[1] == out.shape[1]
            tgt_mask = given
        out = pos_emb(sd, out)
        for i in range(_n_layers(sd, "decoder_layers")):
            out = fft_block(sd, f"decoder_layers.{i}", out, ~tgt_mask, n_head)
        logits = out @ sd["head.weight"].t() + sd["head.bias"]
    return dict(logits=logits, src_mask=src_mask, tgt_mask=tgt_mask, log_dur=log_dur, dur=dur,
                out_lens=out_lens, enc_out=enc_out)


def parrot_infer(sd, batch, n_head=2):
    """Parrot.infer (modules/parrot.py:112-120) -> list[list[int]]."""
    r = parrot_forward(sd, batch, n_head)
    ids = torch.argmax(r["logits"], dim=-1)
    return [c[m].numpy().tolist() for c, m in zip(ids, r["tgt_mask"])], r
