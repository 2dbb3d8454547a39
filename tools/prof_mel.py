"""Dev/profiling driver: ONE device log-mel spectrogram of 8 x 327 680 samples (the audio of BASELINE configs[1])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from parrot_tts_b200.vocoder.mel import mel_spectrogram
y = (torch.rand(8, 327680, device="cuda:0") * 2 - 1) * 0.8
for _ in range(2):
    m = mel_spectrogram(y, 1024, 80, 16000, 256, 1024, 0, 8000)
torch.cuda.synchronize()
print("done", tuple(m.shape), float(m.mean()))
