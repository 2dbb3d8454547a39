"""Dev: wall-clock breakdown of the single-utterance path (cfg1: 64 chars -> 512 units -> PCM on host)."""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200 import configs, ops, synth
from parrot_tts_b200.tte import Parrot
from parrot_tts_b200.vocoder import CodeGenerator, GraphedGenerator
dev = "cuda:0"
h = configs.default_vocoder_config()
voc = CodeGenerator(h); voc.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0)); voc = voc.to(dev)
d = tempfile.mkdtemp(); json.dump({f"s{i}": i for i in range(10)}, open(os.path.join(d, "speakers.json"), "w"))
cfg = configs.default_tte_config(root_path=d)
tte = Parrot(cfg, 150, 0); tte.load_state_dict(synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0, dur_mode="const8")); tte = tte.eval().to(dev)
ph = torch.randint(2, 150, (1, 64)).to(dev); spk = torch.zeros(1, dtype=torch.long, device=dev); sp1 = spk.view(1, 1)
batch = {"phones": ph, "src_mask": ph != 0, "speaker": spk}
ids, _ = tte.infer_device(batch)
gg = GraphedGenerator(voc, 1, ids.shape[1], pcm16=True)
def wall(fn, n=30):
    for _ in range(3): fn()
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
def gpu(fn, n=30):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("tte.infer_device      wall %.3f ms   back-to-back %.3f ms" % (wall(lambda: tte.infer_device(batch)), gpu(lambda: tte.infer_device(batch))))
print("tte._encode only      wall %.3f ms" % wall(lambda: tte._encode(batch)))
print("voc forward           wall %.3f ms   back-to-back %.3f ms" % (wall(lambda: voc(code=ids, spkr=sp1)), gpu(lambda: voc(code=ids, spkr=sp1))))
print("voc graph (+pcm)      wall %.3f ms   back-to-back %.3f ms" % (wall(lambda: gg(ids, sp1)), gpu(lambda: gg(ids, sp1))))
w = voc(code=ids, spkr=sp1)
print("pcm + d2h             wall %.3f ms" % wall(lambda: ops.wav_to_pcm16(w).cpu()))
print("launches tte %d voc %d" % (tte.last_launches, voc.last_launches))
