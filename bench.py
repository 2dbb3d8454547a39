#!/usr/bin/env python
"""bench.py - throughput of the Parrot-TTS inference hot path on B200 (contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]

Workload (default, BASELINE.json configs[1]): batch = 8 utterances x 1024 HuBERT units through the HiFi-GAN
CodeGenerator -> 8 x 327 680 samples per step per GPU.  `--workload cfg3` is configs[2]: 32 x 256 chars ->
TTE -> 32 x 2048 units -> CodeGenerator (text -> wav); the default run also reports it under "text2wav".

One JSON line on stdout (rank 0).  `value` = audio samples/s with the inputs resident in HBM; `e2e` = the same
metric through the public API (vocoder.generate(): pinned host codes -> H2D -> forward -> int16 PCM -> D2H);
`roofline` = the ResBlock conv1d kernel family (90 launches/step) against the measured HBM copy bandwidth;
`cpu_baseline` = the torch-CPU oracle port of the reference on this box's host cores (bounded sample).
`--impl reference` times only that CPU port (the reference is pure PyTorch; its modules cannot travel to the
GPU box, the oracle restates them and is pinned to them by tests/golden).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HOP = 320
VOC_FLOP_PER_UNIT = 321656832.0
# algorithmic fp32 bytes per unit of the 90 ResBlock convs, by stage
RESBLOCK_BYTES_PER_UNIT = [230400.0, 460800.0, 921600.0, 921600.0, 921600.0]
VOC_BYTES_PER_UNIT = 3616000.0
METRIC = "audio samples/s (22.05 kHz) end-to-end text->wav; RTF at 1/2/4/8 GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-text2wav", action="store_true")
    ap.add_argument("--no-streaming", action="store_true")
    return ap.parse_args()


def peaks():
    """(HBM GB/s, bf16 TFLOP/s sustained, source).  The driver-written file is read tolerantly (keys may sit one
    level down, a value may be missing): a malformed file must not take the bench line with it."""
    fb_hbm, fb_tf = 6650.0, 1400.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if not os.path.isfile(p):
        return fb_hbm, fb_tf, "fallback (B200_PROFILING.md)"

    def find(d, key):
        if isinstance(d, dict):
            if isinstance(d.get(key), (int, float)) and d[key] > 0:
                return float(d[key])
            for v in d.values():
                r = find(v, key)
                if r is not None:
                    return r
        return None

    try:
        with open(p) as f:
            d = json.load(f)
    except (OSError, ValueError):
        return fb_hbm, fb_tf, "fallback (MEASURED_PEAKS.json unreadable)"
    hbm = find(d, "hbm_gbs")
    tf = find(d, "bf16_tflops_sustained") or find(d, "bf16_tflops")
    if hbm is None and tf is None:
        return fb_hbm, fb_tf, "fallback (MEASURED_PEAKS.json holds neither hbm_gbs nor bf16_tflops)"
    src = "measured (MEASURED_PEAKS.json)"
    if hbm is None or tf is None:
        src += "; fallback for " + ("hbm_gbs" if hbm is None else "bf16_tflops")
    return hbm or fb_hbm, tf or fb_tf, src


def workload_shape(name):
    if name == "cfg2":
        return dict(workload="cfg2: batch=8 utterances x 1024 units, HiFi-GAN CodeGenerator only", B=8, T=1024, S=0)
    return dict(workload="cfg3: batch=32, 256 chars -> 2048 units, TTE + CodeGenerator", B=32, T=2048, S=256)


# --------------------------------------------------------------------------------------------------------
# clocks sampler (NVML in a thread; falls back to nvidia-smi)
# --------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _poll(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                else:
                    import subprocess
                    out = subprocess.run(
                        ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                         "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                    self.samples.append(int(out[0]))
                    self.max_mhz = int(out[1])
                    for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], out[2:]):
                        if v.strip() == "Active":
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02 if nv is not None else 0.2)

    def __enter__(self):
        self._thr = threading.Thread(target=self._poll, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=5)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference (torch CPU, all host threads)
# --------------------------------------------------------------------------------------------------------
def cpu_models():
    import torch
    from parrot_tts_b200 import configs, synth
    h = configs.default_vocoder_config()
    sd_v = synth.synth_vocoder_state_dict(h, seed=0)
    # fold weight norm once, like the reference's inference script does (remove_weight_norm, inference.py:136-137)
    from oracle import voc_oracle
    folded = {}
    for k in list(sd_v):
        if k.endswith(".weight_v"):
            base = k[:-len(".weight_v")]
            folded[base + ".weight"] = voc_oracle.folded_weight(sd_v, base)
        elif not k.endswith(".weight_g"):
            folded[k] = sd_v[k]
    cfg = configs.default_tte_config()
    sd_t = synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0, dur_mode="const8")
    torch.set_num_threads(os.cpu_count() or 1)
    return h, folded, sd_t


def cpu_step(h, sd_v, sd_t, S, B, T, seed):
    """One bounded CPU sample: B utterances of T units (from S chars when S > 0) -> samples produced, seconds."""
    import torch
    from oracle import tte_oracle, voc_oracle
    g = torch.Generator().manual_seed(seed)
    t0 = time.perf_counter()
    if S > 0:
        phones = torch.randint(2, 150, (B, S), generator=g)
        batch = {"phones": phones, "src_mask": phones != 0, "speaker": torch.arange(B) % 10}
        ids, r = tte_oracle.parrot_infer(sd_t, batch)
        code = torch.argmax(r["logits"], -1)
    else:
        code = torch.randint(0, 1000, (B, T), generator=g)
    spkr = torch.randint(0, 10, (B, 1), generator=g)
    wav = voc_oracle.code_generator_forward(sd_v, h, code, spkr)
    pcm = voc_oracle.to_int16(wav)
    return int(pcm.size), time.perf_counter() - t0


def cpu_sample_shape(shape):
    # bounded sample of the same workload: one utterance of the workload's length
    if shape["S"] > 0:
        return dict(B=1, S=shape["S"], T=shape["T"])
    return dict(B=1, S=0, T=shape["T"])


def cpu_baseline_sample(shape, budget_s=10.0, max_steps=64):
    """The `cpu_baseline` object: the oracle port on this box's host cores over a bounded sample of the workload - one-utterance
    steps of the workload's length until ~``budget_s`` seconds of CPU work have been timed (mean over all of them)."""
    import torch
    smp = cpu_sample_shape(shape)
    hh, sd_v, sd_t = cpu_models()
    cpu_step(hh, sd_v, sd_t, smp["S"], 1, min(smp["T"], 64), 99)  # warm-up
    tot_n, tot_s, best, done = 0, 0.0, None, 0
    while done < max_steps and tot_s < budget_s:
        n, sec = cpu_step(hh, sd_v, sd_t, smp["S"], smp["B"], smp["T"], done)
        tot_n += n
        tot_s += sec
        best = n / sec if best is None else max(best, n / sec)
        done += 1
    return {"value": tot_n / tot_s, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port", "best_step": best,
            "sample": f"{done} step(s) of {smp['B']} utterance x {smp['T']} units ({'TTE+' if smp['S'] else ''}CodeGenerator) = "
                      f"{tot_s:.1f} s of CPU work, mean; torch CPU fp32 oracle port"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    shape = workload_shape(args.workload)
    smp = cpu_sample_shape(shape)
    h, sd_v, sd_t = cpu_models()
    for i in range(max(args.warmup, 1)):
        cpu_step(h, sd_v, sd_t, smp["S"], 1, min(smp["T"], 64) if i == 0 else smp["T"], 100 + i)
    tot_s, tot_n = 0.0, 0
    for i in range(args.steps):
        n, s = cpu_step(h, sd_v, sd_t, smp["S"], smp["B"], smp["T"], i)
        tot_n += n
        tot_s += s
    v = tot_n / tot_s
    sample = f"per step: {smp['B']} utterance x {smp['T']} units of the workload ({'TTE+' if smp['S'] else ''}CodeGenerator, fp32, torch {torch.__version__} CPU)"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": tot_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded random weights and unit codes)",
        "config": {"workload": shape["workload"], "sample": sample},
        "rtf_16k": v / 16000.0, "rtf_22k05": v / 22050.0,
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------------
def run_ours(args):
    import tempfile

    import torch
    import torch.distributed as dist

    from parrot_tts_b200 import _lib, configs, sharding, synth
    from parrot_tts_b200.tte import Parrot
    from parrot_tts_b200.vocoder import CodeGenerator, generate

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback; use --impl reference for the CPU arm)")
    _lib.load()  # fails loudly if the extension is missing
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    quiet = _StdoutToStderr() if world > 1 else None     # NCCL prints its version banner to stdout
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = args.steps, max(args.warmup, 3)
    shape = workload_shape(args.workload)
    B, T, S = shape["B"], shape["T"], shape["S"]
    hbm_peak, tf_peak, peak_src = peaks()

    # ---- weights: rank 0 builds them, ONE broadcast at init, no collective afterwards ---------------------
    h = configs.default_vocoder_config()
    d = tempfile.mkdtemp(prefix="parrot_bench_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"s{i}": i for i in range(10)}, f)
    max_len = 3500
    cfg = configs.default_tte_config(root_path=d, max_len=max_len)
    sd = None
    if rank == 0:
        sd = {"voc." + k: v for k, v in synth.synth_vocoder_state_dict(h, seed=0).items()}
        sd.update({"tte." + k: v for k, v in synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0, dur_mode="const8").items()})
    if world > 1:
        sd = sharding.broadcast_state_dict(sd, src=0, device=dev)
    voc = CodeGenerator(h)
    voc.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("voc.")})
    voc = voc.to(dev)
    tte = Parrot(cfg, 150, 0)
    tte.load_state_dict({k[4:]: v for k, v in sd.items() if k.startswith("tte.")})
    tte = tte.eval().to(dev)
    del sd

    g = torch.Generator().manual_seed(1000 + rank)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, n_warm, n_steps, sampler=None):
        """n_steps calls of fn, each between its own CUDA events on the current stream, L2 flushed before each
        (outside the events); returns mean ms per step, max over ranks."""
        for _ in range(n_warm):
            fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
        ctx = sampler if sampler is not None else _Null()
        with ctx:
            for a, b in evs:
                flush.fill_(1)
                a.record()
                fn()
                b.record()
            barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs) / n_steps
        return max_over_ranks(ms)

    # ---- device-resident inputs ------------------------------------------------------------------------
    code = torch.randint(0, 1000, (B, T), generator=g).to(dev)
    spkr = torch.randint(0, 10, (B, 1), generator=g).to(dev)
    phones = torch.randint(2, 150, (B, max(S, 1)), generator=g).to(dev)
    tbatch = {"phones": phones, "src_mask": phones != 0, "speaker": (torch.arange(B) % 10).to(dev)}
    out = {}

    def step_voc():
        out["wav"] = voc(code=code, spkr=spkr)

    def step_full():
        ids, _ = tte.infer_device(tbatch)
        out["wav"] = voc(code=ids, spkr=spkr)
        out["T"] = ids.shape[1]

    step = step_full if S > 0 else step_voc
    step()
    torch.cuda.synchronize()
    launches_per_step = voc.last_launches + (tte.last_launches if S > 0 else 0)
    samples_per_step = out["wav"].numel()
    assert samples_per_step == B * T * HOP, (samples_per_step, B, T)

    # section events on for the timed run (13 cudaEventRecord per forward)
    voc.set_timing(True, dev)
    sampler = ClockSampler(local)
    # the section events are reused by every forward, so they are read right after each step's end event
    # (reading synchronises; each step is timed by its own event pair, so this does not enter the timing)
    sec_acc = {}
    for _ in range(warmup):
        step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with sampler:
        for a, b in evs:
            flush.fill_(1)
            a.record()
            step()
            b.record()
            for k, v in voc.section_ms().items():
                sec_acc[k] = sec_acc.get(k, 0.0) + v
        barrier()
    ms_step = max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / steps)
    voc.set_timing(False, dev)
    value = world * samples_per_step / (ms_step * 1e-3)
    sec_ms = {k: v / steps for k, v in sec_acc.items()}

    # ---- roofline of the dominant kernel family: conv1d_kernel on the ResBlock path ------------------------
    units = B * T
    rb_ms = sum(sec_ms[f"resblocks{i}"] for i in range(5))
    rb_bytes = sum(RESBLOCK_BYTES_PER_UNIT) * units
    n_rb_launch = 72      # 18 x 3 tc_conv launches (stages 1-3) + 9 x 2 fused tc_pair launches (stages 4-5) for the 90 ResBlock convs
    achieved = rb_bytes / (rb_ms * 1e-3) / 1e9
    # per-unit MMACs of the 90 ResBlock convs: 3 products per MAC on the tensor pipe (split-bf16, 2 MMAs
    # per k-step with the stacked operand for N <= 128, 3 for N = 256)
    rb_mmac_per_unit = 41.288e6 * 3 + 20.644e6 + 10.322e6
    roofline = {
        "bound": "hbm",
        "kernel": "tc_conv_kernel + tc_pair_kernel (the 90 ResBlock1 dilated convs as tcgen05 implicit GEMMs on split-bf16 operands: "
                  "54 single-conv launches for C >= 64, 18 fused conv1->conv2 launches for C = 32 / 16; lrelu + bias + residual + "
                  "resblock mean fused in the epilogues)",
        "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": None,
        "peak_source": peak_src,
        "bytes_per_launch_avg": rb_bytes / n_rb_launch, "ms_per_launch_avg": rb_ms / n_rb_launch,
        "share_of_step": rb_ms / ms_step,
        "algorithmic_tflops": 2 * rb_mmac_per_unit * units / (rb_ms * 1e-3) / 1e12,
        "tensor_pipe_tflops_3pass": 2 * 3 * rb_mmac_per_unit * units / (rb_ms * 1e-3) / 1e12,
        "tensor_pipe_frac_of_bf16_peak": 2 * 3 * rb_mmac_per_unit * units / (rb_ms * 1e-3) / 1e12 / tf_peak,
        "per_stage": {f"stage{i + 1}": {"ms": sec_ms[f"resblocks{i}"], "GB/s": RESBLOCK_BYTES_PER_UNIT[i] * units / (sec_ms[f"resblocks{i}"] * 1e-3) / 1e9,
                                        "frac": RESBLOCK_BYTES_PER_UNIT[i] * units / (sec_ms[f"resblocks{i}"] * 1e-3) / 1e9 / hbm_peak}
                      for i in range(5)},
        "section_ms": sec_ms,
    }
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tr):
        try:
            with open(tr) as f:
                roofline["traffic"] = json.load(f).get(args.workload)
        except Exception:
            pass

    # ---- e2e through the public API: pinned host codes -> H2D -> generate() -> int16 PCM on the host ------
    code_h = code.cpu().pin_memory()
    spkr_h = spkr.cpu().pin_memory()
    phones_h = phones.cpu().pin_memory()
    spk_ids_h = (torch.arange(B) % 10).pin_memory()
    res = {}

    def step_e2e():
        if S > 0:
            ph = phones_h.to(dev, non_blocking=True)
            b = {"phones": ph, "src_mask": ph != 0, "speaker": spk_ids_h.to(dev, non_blocking=True)}
            ids, _ = tte.infer_device(b)
            c = {"code": ids, "spkr": spkr_h.to(dev, non_blocking=True)}
        else:
            c = {"code": code_h.to(dev, non_blocking=True), "spkr": spkr_h.to(dev, non_blocking=True)}
        audio, _ = generate(h, voc, c)      # int16 numpy on the host
        res["pcm"] = audio

    e2e_ms = timed(step_e2e, warmup, steps)
    h2d = (phones_h.numel() * 8 + spk_ids_h.numel() * 8 if S > 0 else code_h.numel() * 8) + spkr_h.numel() * 8
    d2h = res["pcm"].size * 2
    e2e = {"value": world * samples_per_step / (e2e_ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms,
           "api": "parrot_tts_b200.vocoder.generate(h, CodeGenerator, code) -> int16 PCM on host"}

    # ---- supplementary: configs[2] text -> wav (TTE + vocoder) ----------------------------------------------
    text2wav = None
    if S == 0 and not args.no_text2wav:
        B3, S3 = 32, 256
        ph3 = torch.randint(2, 150, (B3, S3), generator=g).to(dev)
        b3 = {"phones": ph3, "src_mask": ph3 != 0, "speaker": (torch.arange(B3) % 10).to(dev)}
        sp3 = (torch.arange(B3) % 10).view(B3, 1).to(dev)
        t3 = {}

        def step3():
            ids, _ = tte.infer_device(b3)
            t3["n"] = voc(code=ids, spkr=sp3).numel()
            t3["launch"] = tte.last_launches + voc.last_launches

        def step3_tte():
            tte.infer_device(b3)

        ms3 = timed(step3, 2, 3)
        ms3_tte = timed(step3_tte, 1, 3)
        tte_flop = 2.45e12   # TTE at cfg3
        text2wav = {"workload": "cfg3: batch=32, 256 chars -> 2048 units, TTE (tcgen05 split-fp16 GEMMs + fused attention) + CodeGenerator",
                    "ms_per_step": ms3, "tte_algorithmic_tflops": tte_flop / (ms3_tte * 1e-3) / 1e12,
                    "tte_ms": ms3_tte, "value": world * t3["n"] / (ms3 * 1e-3), "unit": "samples/s",
                    "rtf_16k": t3["n"] / (ms3 * 1e-3) / 16000.0, "gpu_launches_per_step": t3["launch"]}

    # ---- supplementary: configs[0] single-utterance latency: 64 chars -> 512 units -> int16 PCM on the host ---------
    latency = None
    if S == 0 and not args.no_text2wav:
        from parrot_tts_b200 import ops as _ops
        ph1_h = torch.randint(2, 150, (1, 64), generator=g).pin_memory()
        sp1 = torch.zeros(1, 1, dtype=torch.long, device=dev)
        spk1 = torch.zeros(1, dtype=torch.long, device=dev)

        from parrot_tts_b200.vocoder import GraphedGenerator
        voc_graphs = {}

        def run1(graphed=False):
            ph = ph1_h.to(dev, non_blocking=True)
            ids, _ = tte.infer_device({"phones": ph, "src_mask": ph != 0, "speaker": spk1})
            if not graphed:
                return _ops.wav_to_pcm16(voc(code=ids, spkr=sp1)).cpu().numel()
            T1 = int(ids.shape[1])
            if T1 not in voc_graphs:      # the vocoder forward + PCM cast of this length as one CUDA graph
                voc_graphs[T1] = GraphedGenerator(voc, 1, T1, pcm16=True)
            return voc_graphs[T1](ids, sp1).cpu().numel()

        def wall(fn):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(20):
                torch.cuda.synchronize()
                t0 = time.perf_counter(); n = fn(); ts.append((time.perf_counter() - t0) * 1e3)
            ts.sort()
            return ts, n

        ts, n1 = wall(run1)
        latency = {"workload": "cfg1: 1 utterance, 64 chars -> TTE -> 512 units -> CodeGenerator -> int16 PCM on host (wall clock)",
                   "median_ms": ts[len(ts) // 2], "min_ms": ts[0], "samples": n1, "rtf_16k": n1 / (ts[len(ts) // 2] * 1e-3) / 16000.0,
                   "gpu_launches": tte.last_launches + voc.last_launches + 1}
        try:      # supplementary: never let a capture problem take the headline line down
            tg, _ = wall(lambda: run1(True))
            latency["vocoder_cuda_graph"] = {"median_ms": tg[len(tg) // 2], "min_ms": tg[0]}
        except Exception as e:  # noqa: BLE001
            latency["vocoder_cuda_graph"] = {"error": str(e)[:200]}

    # ---- supplementary: configs[4] long-form streaming: 16 384 units in 1024-unit chunks with a 32-unit halo -------
    streaming = None
    if S == 0 and not args.no_streaming:
        from parrot_tts_b200.vocoder import stream_generate
        code5_h = torch.randint(0, 1000, (1, 16384), generator=g).pin_memory()
        spk5 = torch.zeros(1, 1, dtype=torch.long, device=dev)

        stream_graphs = {}

        def run_stream(first_only, graphs=None):
            c5 = code5_h.to(dev, non_blocking=True)
            n = 0
            for _, pcm in stream_generate(voc, c5, spk5, chunk_units=1024, halo_units=32, pcm16=True, graphs=graphs):
                n += pcm.cpu().numel()       # int16 PCM of the chunk on the host
                if first_only:
                    break
            return n

        run_stream(False)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); run_stream(True); first_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); n5 = run_stream(False); all_s = time.perf_counter() - t0
        streaming = {"workload": "cfg5: 16 384 units streamed as 16 chunks of 1024 units (+32-unit halo), host codes -> int16 PCM on host",
                     "time_to_first_chunk_ms": first_ms, "steady_samples_per_s": n5 / all_s, "rtf_16k": n5 / all_s / 16000.0,
                     "total_ms": all_s * 1e3}
        try:
            run_stream(False, stream_graphs)           # captures one CUDA graph per window length (3)
            torch.cuda.synchronize()
            t0 = time.perf_counter(); run_stream(True, stream_graphs); first_g = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter(); n5g = run_stream(False, stream_graphs); all_g = time.perf_counter() - t0
            streaming["cuda_graphs"] = {"time_to_first_chunk_ms": first_g, "steady_samples_per_s": n5g / all_g, "total_ms": all_g * 1e3}
        except Exception as e:  # noqa: BLE001
            streaming["cuda_graphs"] = {"error": str(e)[:200]}

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload -----------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_sample(shape)

    if rank == 0:
        clocks = sampler.summary()
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (split-bf16 x3 products on tcgen05, fp32 accumulation in TMEM)",
            "data": "synthetic (seeded random weights, random unit codes / char tokens)",
            "config": {"workload": shape["workload"], "units_per_gpu_per_step": units, "samples_per_gpu_per_step": samples_per_step,
                       "l2": "256 MB flush write before every timed step; per-step activation working set (>5 GB) exceeds the 126 MB L2",
                       "sharding": f"{world} rank(s), one utterance batch per rank, weights broadcast once at init, no per-step collective"},
            "rtf_16k": value / 16000.0, "rtf_22k05": value / 22050.0,
            "vocoder_algorithmic_tflops": VOC_FLOP_PER_UNIT * units * world / (ms_step * 1e-3) / 1e12 if S == 0 else None,
            "e2e": e2e, "gpu_launches": int(launches_per_step * steps), "gpu_launches_per_step": int(launches_per_step),
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "text2wav": text2wav, "latency": latency, "streaming": streaming,
        }
        if quiet is not None:
            quiet.restore()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _StdoutToStderr:
    """Route fd 1 to stderr while libraries that print to stdout (NCCL's version banner) run; stdout must carry only the JSON line."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def run_cfg4(args):
    """BASELINE configs[3]: 256 utterances of 256..4096 units, sharded data-parallel by longest-processing-time-first on
    their length, batched in length buckets (each item masked to its own length = its batch-of-1 result); no collective on
    the data path.  value = true (unpadded) samples of all ranks / slowest rank's time."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from parrot_tts_b200 import _lib, configs, sharding, synth
    from parrot_tts_b200.vocoder import CodeGenerator

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    quiet = _StdoutToStderr() if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h = configs.default_vocoder_config()
    sd = synth.synth_vocoder_state_dict(h, seed=0) if rank == 0 else None
    if world > 1:
        sd = sharding.broadcast_state_dict(sd, src=0, device=dev)
    voc = CodeGenerator(h); voc.load_state_dict(sd); voc = voc.to(dev)
    lengths = np.random.default_rng(0).integers(256, 4097, 256).tolist()
    bins = sharding.lpt_assign([sharding.utterance_cost(n, with_tte=False) for n in lengths], world)
    batches = sharding.length_buckets(bins[rank], lengths, max_batch=8)
    g = torch.Generator().manual_seed(7)
    work = []
    for idx in batches:
        Lmax = max(lengths[i] for i in idx)
        code = torch.randint(0, 1000, (len(idx), Lmax), generator=g).to(dev)
        lens = torch.tensor([lengths[i] for i in idx], dtype=torch.int32, device=dev)
        spk = torch.randint(0, 10, (len(idx), 1), generator=g).to(dev)
        work.append((code, spk, lens))
    my_units = sum(lengths[i] for i in bins[rank]); padded = sum(c.numel() for c, _, _ in work)

    def step():
        for code, spk, lens in work:
            voc(code=code, spkr=spk, lens=lens)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    total_units = sum(lengths)
    if rank == 0:
        v = total_units * HOP / (ms * 1e-3)
        if quiet is not None:
            quiet.restore()
        print(json.dumps({"metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f32 (split-bf16 x3 products on tcgen05, fp32 accumulation in TMEM)", "data": "synthetic",
                          "config": {"workload": "cfg4: 256 utterances, 256..4096 units each (numpy default_rng(0)), CodeGenerator, LPT-sharded over the ranks, length buckets of <= 8",
                                     "rank0_units": my_units, "rank0_padded_units": padded, "rank0_batches": len(work)},
                          "rtf_16k": v / 16000.0, "gpu_launches": int(voc.last_launches * len(work) * args.steps)}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg4":
        run_cfg4(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
