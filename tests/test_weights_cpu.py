"""Host logic: the cached weight stamp (parrot_tts_b200/_weights.py) must see every way the weights can change, because the
device copy (folded + packed by *_finalize) is rebuilt only when the stamp differs."""
import json
import os
import tempfile

import torch

from parrot_tts_b200 import configs, synth
from parrot_tts_b200.tte import Parrot
from parrot_tts_b200.vocoder import CodeGenerator


def test_vocoder_stamp_sees_every_kind_of_weight_change():
    h = configs.default_vocoder_config()
    gen = CodeGenerator(h)
    sd = synth.synth_vocoder_state_dict(h, seed=0)
    gen.load_state_dict(sd)
    s0 = gen._stamp("cuda:0")
    assert gen._stamp("cuda:0") == s0                         # stable while nothing changes
    assert gen._stamp("cuda:1") != s0                         # device is part of it
    gen.load_state_dict(sd)                                   # in-place copy_: version counters
    s1 = gen._stamp("cuda:0")
    assert s1 != s0
    with torch.no_grad():
        gen.conv_post.bias.add_(1.0)                          # in-place edit of one tensor
    s2 = gen._stamp("cuda:0")
    assert s2 != s1
    gen.remove_weight_norm()                                  # parameters re-registered: the tensor list is rebuilt
    s3 = gen._stamp("cuda:0")
    assert s3 != s2 and len(s3) < len(s2)
    gen.double()                                              # .to()/.double(): new storage
    s4 = gen._stamp("cuda:0")
    assert s4 != s3
    gen.conv_mode = 0
    assert gen._stamp("cuda:0") != s4                         # kernel family is part of it


def test_tte_stamp_tracks_parameters_and_buffers():
    d = tempfile.mkdtemp(prefix="parrot_tte_")
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"s{i}": i for i in range(10)}, f)
    m = Parrot(configs.default_tte_config(root_path=d), 150, 0)
    a = m._wstamp.stamp()
    assert m._wstamp.stamp() == a
    assert set(m._wstamp.tensors()) == set(m.state_dict())
    with torch.no_grad():
        m.pos_emb.pe.mul_(1.0)                                # a buffer, edited in place
    b = m._wstamp.stamp()
    assert b != a
    m.head.weight = torch.nn.Parameter(torch.zeros_like(m.head.weight))      # re-registration
    c = m._wstamp.stamp()
    assert c != b
    assert m._wstamp.tensors()["head.weight"] is m.head.weight


def test_unrelated_registrations_do_not_stale_an_engine_and_invalidate_does():
    """the process-wide registration counter only refreshes the cached tensor list; it is not part of the comparison key, so
    building another model (or any nn.Module) leaves an existing engine's stamp unchanged.  `.data` edits are invisible to torch's
    version counters: `invalidate_weights()` is the explicit way."""
    h = configs.default_vocoder_config()
    gen = CodeGenerator(h)
    s0 = gen._stamp("cuda:0")
    torch.nn.Linear(4, 4)                                     # unrelated registrations anywhere in the process
    CodeGenerator(h)
    assert gen._stamp("cuda:0") == s0
    gen.conv_post.bias.data.add_(1.0)                         # not seen ...
    assert gen._stamp("cuda:0") == s0
    gen.invalidate_weights()                                  # ... unless said
    assert gen._stamp("cuda:0") != s0
