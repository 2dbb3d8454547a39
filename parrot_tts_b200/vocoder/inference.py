"""The reference's vocoder inference helpers that sit on the hot path (reference utils/vocoder/inference.py:65-74,
104-109,136-137), on the sm_100a backend."""
import time

import torch

from .. import ops

MAX_WAV_VALUE = 32768.0  # reference utils/vocoder/dataset.py:22


def load_checkpoint(filepath, device="cpu"):
    """torch.load of a ``g_XXXXXXXX`` file -> dict with key 'generator' (reference utils/vocoder/inference.py:48-53)."""
    return torch.load(filepath, map_location=device, weights_only=True)


def generate(h, generator, code):
    """(int16 numpy audio, rtf) for ``generator(**code)`` - reference utils/vocoder/inference.py:65-74.

    Same outputs as the reference; the x32768 and the int16 cast (truncation toward zero, two's-complement wrap)
    run on the device so only 2 bytes per sample cross PCIe, and the wall clock brackets a synchronised region
    (the reference's timing omits ``cuda.synchronize`` and under-reports GPU time).
    """
    start = time.time()
    if hasattr(generator, "forward_pcm16"):          # the cast is conv_post's epilogue: no f32 waveform round trip through HBM
        pcm = generator.forward_pcm16(**code)
    else:
        y_g_hat = generator(**code)
        if type(y_g_hat) is tuple:
            y_g_hat = y_g_hat[0]
        pcm = ops.wav_to_pcm16(y_g_hat)
    audio = ops.to_host(pcm.squeeze())
    if hasattr(generator, "check_ids"):
        generator.check_ids()                          # IndexError for ids outside the embedding tables, like nn.Embedding
    rtf = (time.time() - start) / (pcm.shape[-1] / h.sampling_rate)
    return audio, rtf


def normalize(audio, device="cuda"):
    """Drop-in for ``librosa.util.normalize(audio.astype(np.float32))`` as the reference calls it on generated audio
    (utils/vocoder/inference.py:169, demo.ipynb cell 13): int16 PCM -> float32 divided by the utterance's peak.
    numpy in -> numpy out (the reduction and the division run on the device); a device tensor [n] / [B, n] stays there."""
    if torch.is_tensor(audio):
        return ops.peak_normalize(audio)
    import numpy as np
    a = np.ascontiguousarray(audio)
    if a.dtype != np.int16:
        raise TypeError("normalize expects the int16 audio that generate() returns")
    return ops.to_host(ops.peak_normalize(torch.from_numpy(a).to(device)))


class GraphedGenerator:
    """CUDA-graph replay of ``generator(code=, spkr=, lens=)`` for one fixed ``[B, T]`` shape: the ~80 launches of a forward are
    captured once (stream capture keeps the programmatic-dependent-launch edges and the cluster launches) and replayed with
    one driver call, which is what bounds the small, latency-critical calls (a single utterance, the chunks of a stream).
    Inputs are copied into static buffers; the returned waveform (or int16 PCM with ``pcm16=True``) is a static buffer too,
    valid until the next call.  The graph records raw pointers into the engine's weight arena and workspace: every call compares
    the engine handle and its memory epoch (``pb_voc_memory_epoch``) with the ones recorded at capture and captures again when the
    weights were re-uploaded or a larger forward regrew the workspace (capturing the LARGEST shape first avoids the re-capture)."""

    def __init__(self, generator, B, T, pcm16=False):
        dev = next(generator.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedGenerator needs a generator on a CUDA device")
        self.generator, self.B, self.T, self.pcm16 = generator, int(B), int(T), bool(pcm16)
        self.device = dev
        self.code = torch.zeros(B, T, dtype=torch.long, device=dev)
        self.spkr = torch.zeros(B, 1, dtype=torch.long, device=dev) if getattr(generator, "multispkr", None) else None
        self.lens = torch.full((B,), T, dtype=torch.int32, device=dev)
        self.captures = 0
        self._capture()

    def _run(self):
        kw = {"code": self.code, "lens": self.lens}
        if self.spkr is not None:
            kw["spkr"] = self.spkr
        return self.generator.forward_pcm16(**kw) if self.pcm16 else self.generator(**kw)

    def _memory_key(self):
        """(engine handle, its memory epoch): the graph holds raw pointers into the handle's weight arena and workspace."""
        from .. import _lib
        hnd = self.generator._engine(self.device)          # rebuilds the engine if the module's weights changed
        return (hnd.value, int(_lib.load().pb_voc_memory_epoch(hnd)))

    def _capture(self):
        dev = self.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                  # warm-up off the capture: engine, workspace, function attributes
            self._run(); self._run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()
        self._key = self._memory_key()
        self.captures += 1

    def __call__(self, code, spkr=None, lens=None):
        if tuple(code.shape) != (self.B, self.T):
            raise ValueError(f"GraphedGenerator captured for {(self.B, self.T)}, got {tuple(code.shape)}")
        if self._memory_key() != self._key:
            # the weights were re-uploaded (new engine) or a larger forward regrew the workspace since the capture: the recorded
            # pointers are dead - capture again rather than replay into freed memory
            self._capture()
        self.code.copy_(code, non_blocking=True)
        if self.spkr is not None:
            if spkr is None:
                raise KeyError("spkr")
            self.spkr.copy_(spkr.reshape(self.B, 1), non_blocking=True)
        if lens is None:
            self.lens.fill_(self.T)
        else:
            self.lens.copy_(torch.as_tensor(lens).to(torch.int32), non_blocking=True)
        self.graph.replay()
        return self.out


def stream_chunks(n_units, chunk_units=1024, halo_units=32):
    """[(lo, hi, a, b)]: chunk i synthesises units [a, b) from the window [lo, hi) = [a - halo, b + halo) clipped to the
    utterance.  The generator is purely convolutional with a receptive-field radius of ~23 units,
    so with halo >= 24 every kept sample equals the unchunked result."""
    out = []
    for a in range(0, n_units, chunk_units):
        b = min(a + chunk_units, n_units)
        out.append((max(a - halo_units, 0), min(b + halo_units, n_units), a, b))
    return out


def stream_generate(generator, code, spkr=None, chunk_units=1024, halo_units=32, pcm16=False, graphs=None, lookahead=0, host=False):
    """Long-form synthesis in chunks (BASELINE configs[4]): yields (first_sample, audio) per chunk, audio f32 [B, n] on the
    device (or int16 with ``pcm16=True``), n = hop * (b - a).  ``code`` i64 [B, T] on the device.  Each chunk is one
    ``generator(code=window, spkr=spkr)`` call on the window of units it needs; nothing is recomputed beyond the two halos.
    ``host=True`` yields numpy arrays: every chunk is copied to pinned host memory on a copy stream as soon as its own kernels are
    done.  ``lookahead=k`` launches k further chunks before a chunk is handed over, so that the copy (and the consumer) of chunk i
    overlap the kernels of chunk i + 1 - steady-state throughput; the first chunk then arrives k launches later.
    ``graphs={}`` replays one CUDA graph per distinct window length instead (the dict can be kept for the next stream; a yielded
    chunk is then a view of the graph's static output buffer, valid until the next chunk of the same window length)."""
    T = code.shape[1]
    hop = generator.hop
    chunks = stream_chunks(T, chunk_units, halo_units)
    if graphs is not None:      # dict {(window length, pcm16): GraphedGenerator}, filled here (largest window first), reusable across streams
        for n in sorted({hi - lo for lo, hi, _, _ in chunks}, reverse=True):
            if (n, bool(pcm16)) not in graphs:
                graphs[(n, bool(pcm16))] = GraphedGenerator(generator, code.shape[0], n, pcm16=pcm16)
    if lookahead and graphs is not None:
        raise ValueError("lookahead needs the plain path: a graph's output buffer is overwritten by the next chunk of the same window length")
    held, copy_stream = [], None
    for lo, hi, a, b in chunks:
        if graphs is not None:
            out = graphs[(hi - lo, bool(pcm16))](code[:, lo:hi], spkr)
        else:
            kw = {"code": code[:, lo:hi].contiguous()}
            if spkr is not None:
                kw["spkr"] = spkr
            out = generator.forward_pcm16(**kw) if pcm16 else generator(**kw)      # int16 straight from conv_post's epilogue
        piece = out[:, 0, (a - lo) * hop:(b - lo) * hop]
        done = None
        if host:      # the chunk leaves for pinned host memory on a copy stream as soon as ITS kernels are done, whatever is queued behind
            dev = piece.device
            if copy_stream is None:
                copy_stream = torch.cuda.Stream(dev)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(dev))
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                buf = torch.empty(piece.shape, dtype=piece.dtype, pin_memory=True)
                buf.copy_(piece, non_blocking=True)
                piece.record_stream(copy_stream)
                done = torch.cuda.Event()
                done.record(copy_stream)
            piece = buf
        held.append((a * hop, piece, done))
        if len(held) > lookahead:       # hand chunk i over only once chunk i + lookahead is launched: its copy overlaps their kernels
            yield _hand_over(held.pop(0))
    while held:
        yield _hand_over(held.pop(0))


def _hand_over(item):
    first, piece, done = item
    if done is None:
        return first, piece
    done.synchronize()
    return first, piece.numpy()
