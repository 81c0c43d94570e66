#!/usr/bin/env python3
"""Generate the committed golden fixtures (run in the build container, where
/root/reference and oracle/_ref exist):

 * speech_16k.npz / speech_48k.npz : mono excerpts (channel 0) of the
   reference's own test recordings tests/resources/{far,near}{16,48}_stereo.pcm,
   used as realistic inputs on the GPU box where /root/reference is absent.
 * hpf_kat.npz : the known-answer vectors of the reference's own high-pass
   filter unit test (tests/unit/high_pass_filter_unittest.cc:198-330), parsed
   from the literal arrays there.
 * ref_outputs.npz : outputs of the compiled reference (oracle/_ref) for the
   parity configurations, so the oracle itself is pinned against a recorded
   run (detects a silently different oracle build).
"""
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import ref  # noqa: E402

RES = "/root/reference/tests/resources"


def mono(name, n):
    return np.fromfile(os.path.join(RES, name), dtype=np.int16).reshape(-1, 2)[:n, 0].copy()


def main():
    far16, near16 = mono("far16_stereo.pcm", 112000), mono("near16_stereo.pcm", 112000)
    far48, near48 = mono("far48_stereo.pcm", 144000), mono("near48_stereo.pcm", 144000)
    np.savez_compressed(os.path.join(HERE, "speech_16k.npz"), far=far16, near=near16)
    np.savez_compressed(os.path.join(HERE, "speech_48k.npz"), far=far48, near=near48)

    # HPF known-answer test, stereo/mono x rates, from the reference unit test.
    src = open("/root/reference/tests/unit/high_pass_filter_unittest.cc").read()
    kats = {}
    fl = lambda t: np.array([float(x.rstrip("f")) for x in re.findall(r"[-+]?\d+\.\d+f", t)], dtype=np.float32)
    for m in re.finditer(r"TEST\(HighPassFilterAccuracyTest, (Mono\w+)\)\s*\{(.*?)\n\}", src, re.S):
        name, body = m.group(1), m.group(2)
        kin = fl(re.search(r"kReferenceInput\[\] = \{(.*?)\};", body, re.S).group(1))
        kref = fl(re.search(r"kReference\[\] = \{(.*?)\};", body, re.S).group(1))
        assert kin.size % 160 == 0 and kref.size == 12, (name, kin.size, kref.size)
        kats[name + "_in"] = kin
        kats[name + "_ref"] = kref
    assert kats
    np.savez(os.path.join(HERE, "hpf_kat.npz"), **kats)

    outs = {}
    for tag, kw, rate, far, near in (
            ("ns_mod_16k", dict(aec=False, ns=True, ns_level=1), 16000, None, near16),
            ("ns_high_48k", dict(aec=False, ns=True, ns_level=2), 48000, None, near48),
            ("aec_ns_16k", dict(aec=True, ns=True, ns_level=1), 16000, far16, near16),
            ("aec_16k", dict(aec=True, ns=False), 16000, far16, near16)):
        a = ref.RefApm(max_rate=48000, **kw)
        out, stats, err = a.run_i16(rate, far, near, stats_every=100)
        assert err == 0
        outs[tag] = out
        outs[tag + "_stats"] = stats
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **outs)
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
