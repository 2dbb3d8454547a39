#!/bin/bash
# compute-sanitizer on small shapes: memcheck + racecheck over the vocoder (tensor-core path incl. fused pairs) and the TTE
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import json, os, sys, tempfile
sys.path.insert(0, os.getcwd())
import torch
from parrot_tts_b200 import configs, synth
from parrot_tts_b200.vocoder import CodeGenerator
from parrot_tts_b200.tte import Parrot
h = configs.default_vocoder_config()
gen = CodeGenerator(h); gen.load_state_dict(synth.synth_vocoder_state_dict(h, seed=0)); gen = gen.cuda()
code = torch.randint(0, 1000, (2, 9)).cuda(); spkr = torch.randint(0, 10, (2, 1)).cuda()
w = gen(code=code, spkr=spkr, lens=torch.tensor([9, 5]))
d = tempfile.mkdtemp(); json.dump({f"s{i}": i for i in range(10)}, open(os.path.join(d, "speakers.json"), "w"))
cfg = configs.default_tte_config(root_path=d)
m = Parrot(cfg, 150, 0); m.load_state_dict(synth.synth_tte_state_dict(cfg, 150, 0, 10, seed=0)); m = m.eval().cuda()
ph = torch.randint(2, 150, (2, 7)).cuda(); ph[1, 5:] = 0
ids, n = m.infer_device({"phones": ph, "src_mask": ph != 0, "speaker": torch.tensor([0, 3]).cuda()})
torch.cuda.synchronize()
print("done", float(w.abs().max()), ids.shape)
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/san.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/san_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/san.py > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/san_racecheck.log
