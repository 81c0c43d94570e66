#!/usr/bin/env python3
"""Generate webrtc-audio-processing_b200/csrc/wap_tables.inc.

Every table is COMPUTED here from its published definition (Ooura's makewt /
makect recurrences, Hann windows, log / sine tables) in the arithmetic the
reference used to produce its literals; when /root/reference is present the
result is additionally validated bit-for-bit against the literal arrays in the
reference sources (read only for that check).  The generated file is committed.
"""
import math
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "webrtc-audio-processing_b200", "csrc", "wap_tables.inc")
REF = "/root/reference/webrtc"
f32 = np.float32


def bitrv_perm(n):
    """Permutation applied by Ooura's bitrv2(n) to the n/2 complex elements,
    obtained by running its swap schedule on an index array."""
    ip = [0]
    l, m = n, 1
    while (m << 3) < l:
        l >>= 1
        ip = ip + [x + l for x in ip[:m]]
        m <<= 1
    m2 = 2 * m
    a = list(range(n))

    def swap(j1, k1):
        a[j1], a[k1] = a[k1], a[j1]
        a[j1 + 1], a[k1 + 1] = a[k1 + 1], a[j1 + 1]

    if (m << 3) == l:
        for k in range(m):
            for j in range(k):
                j1 = 2 * j + ip[k]
                k1 = 2 * k + ip[j]
                swap(j1, k1)
                j1 += m2; k1 += 2 * m2
                swap(j1, k1)
                j1 += m2; k1 -= m2
                swap(j1, k1)
                j1 += m2; k1 += 2 * m2
                swap(j1, k1)
            j1 = 2 * k + m2 + ip[k]
            swap(j1, j1 + m2)
    else:
        for k in range(1, m):
            for j in range(k):
                j1 = 2 * j + ip[k]
                k1 = 2 * k + ip[j]
                swap(j1, k1)
                swap(j1 + m2, k1 + m2)
    return [a[2 * i] // 2 for i in range(n // 2)], ip


def bitrv_apply(vals, n, ip_unused=None):
    perm, _ = bitrv_perm(n)
    out = list(vals)
    for i in range(n // 2):
        out[2 * i] = vals[2 * perm[i]]
        out[2 * i + 1] = vals[2 * perm[i] + 1]
    return out


def makewt(nw, cosf, sinf):
    nwh = nw >> 1
    delta = f32(f32(math.atan(1.0)) / f32(nwh))
    w = [f32(0)] * nw
    w[0] = f32(1); w[1] = f32(0)
    w[nwh] = cosf(delta * f32(nwh)); w[nwh + 1] = w[nwh]
    for j in range(2, nwh, 2):
        x = cosf(delta * f32(j)); y = sinf(delta * f32(j))
        w[j] = x; w[j + 1] = y; w[nw - j] = y; w[nw - j + 1] = x
    return bitrv_apply(w, nw)


def makect(nc, cosf, sinf):
    nch = nc >> 1
    delta = f32(f32(math.atan(1.0)) / f32(nch))
    c = [f32(0)] * nc
    c[0] = cosf(delta * f32(nch)); c[nch] = f32(0.5) * c[0]
    for j in range(1, nch):
        c[j] = f32(0.5) * cosf(delta * f32(j))
        c[nc - j] = f32(0.5) * sinf(delta * f32(j))
    return c


def twiddles(w, nblocks):
    """Per radix-4 block q: (c1,s1,c2,s2,c3,s3) in the form out = (c*x.r - s*x.i, c*x.i + s*x.r)."""
    tw = []
    for q in range(nblocks):
        k1 = 2 * (q // 2); k2 = 2 * k1
        wk2r, wk2i = w[k1], w[k1 + 1]
        if q % 2 == 0:
            wk1r, wk1i = w[k2], w[k2 + 1]
            wk3r = f32(wk1r - f32(f32(f32(2) * wk2i) * wk1i))
            wk3i = f32(f32(f32(f32(2) * wk2i) * wk1r) - wk1i)
            c2, s2 = wk2r, wk2i
        else:
            wk1r, wk1i = w[k2 + 2], w[k2 + 3]
            wk3r = f32(wk1r - f32(f32(f32(2) * wk2r) * wk1i))
            wk3i = f32(f32(f32(f32(2) * wk2r) * wk1r) - wk1i)
            c2, s2 = f32(-wk2i), wk2r
        tw.append((wk1r, wk1i, c2, s2, wk3r, wk3i))
    return tw


def ref_floats(path, name):
    src = open(os.path.join(REF, path)).read()
    m = re.search(re.escape(name) + r"[^=;]*=\s*\{(.*?)\};", src, re.S)
    assert m, (path, name)
    body = re.sub(r"//.*", "", m.group(1))
    body = body.replace("ln10_v<float>", repr(float(f32(math.log(10.0))))).replace(
        "sqrt2_v<float>", repr(float(f32(math.sqrt(2.0)))))
    return [f32(float(x.rstrip("f"))) for x in re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?f?", body)]


def check(name, mine, path, refname, n=None):
    if not os.path.isdir(REF):
        return
    ref = ref_floats(path, refname)
    if n:
        ref = ref[:n]
    mine = [f32(x) for x in mine]
    bad = [i for i, (a, b) in enumerate(zip(mine, ref)) if a != b and not (a == 0 and b == 0)]
    assert len(mine) == len(ref) and not bad, (name, len(mine), len(ref), bad[:8],
                                              [(float(mine[i]), float(ref[i])) for i in bad[:4]])
    print("validated", name, "against", path)


def flit(v):
    t = "%.9g" % float(v)
    if "." not in t and "e" not in t:
        t += ".0"
    return t + "f"


def farr(name, vals, per=6):
    s = "alignas(16) WAP_DEVCONST float %s[%d] = {\n" % (name, len(vals))
    for i in range(0, len(vals), per):
        s += "    " + ", ".join(flit(v) for v in vals[i:i + per]) + ",\n"
    return s + "};\n"


def iarr(name, vals, ctype="unsigned char", per=16):
    s = "alignas(16) WAP_DEVCONST %s %s[%d] = {\n" % (ctype, name, len(vals))
    for i in range(0, len(vals), per):
        s += "    " + ", ".join(str(int(v)) for v in vals[i:i + per]) + ",\n"
    return s + "};\n"


RDFT_W128 = [
    1.0000000000, 0.0000000000, 0.7071067691, 0.7071067691, 0.9238795638, 0.3826834559, 0.3826834559,
    0.9238795638, 0.9807852507, 0.1950903237, 0.5555702448, 0.8314695954, 0.8314695954, 0.5555702448,
    0.1950903237, 0.9807852507, 0.9951847196, 0.0980171412, 0.6343933344, 0.7730104327, 0.8819212914,
    0.4713967443, 0.2902846634, 0.9569403529, 0.9569403529, 0.2902846634, 0.4713967443, 0.8819212914,
    0.7730104327, 0.6343933344, 0.0980171412, 0.9951847196, 0.7071067691, 0.4993977249, 0.4975923598,
    0.4945882559, 0.4903926253, 0.4850156307, 0.4784701765, 0.4707720280, 0.4619397819, 0.4519946277,
    0.4409606457, 0.4288643003, 0.4157347977, 0.4016037583, 0.3865052164, 0.3704755902, 0.3535533845,
    0.3357794881, 0.3171966672, 0.2978496552, 0.2777851224, 0.2570513785, 0.2356983721, 0.2137775421,
    0.1913417280, 0.1684449315, 0.1451423317, 0.1214900985, 0.0975451618, 0.0733652338, 0.0490085706,
    0.0245338380,
]


NS_LOG_TABLE = [
    0.000000, 0.000000, 0.000000, 0.000000, 0.000000, 1.609438, 1.791759, 1.945910,
    2.079442, 2.197225, 2.302585, 2.397895, 2.484907, 2.564949, 2.639057, 2.708050,
    2.772589, 2.833213, 2.890372, 2.944439, 2.995732, 3.044522, 3.091043, 3.135494,
    3.178054, 3.218876, 3.258097, 3.295837, 3.332205, 3.367296, 3.401197, 3.433987,
    3.465736, 3.496507, 3.526361, 3.555348, 3.583519, 3.610918, 3.637586, 3.663562,
    3.688879, 3.713572, 3.737669, 3.761200, 3.784190, 3.806663, 3.828641, 3.850147,
    3.871201, 3.891820, 3.912023, 3.931826, 3.951244, 3.970292, 3.988984, 4.007333,
    4.025352, 4.043051, 4.060443, 4.077538, 4.094345, 4.110874, 4.127134, 4.143135,
    4.158883, 4.174387, 4.189655, 4.204693, 4.219508, 4.234107, 4.248495, 4.262680,
    4.276666, 4.290460, 4.304065, 4.317488, 4.330733, 4.343805, 4.356709, 4.369448,
    4.382027, 4.394449, 4.406719, 4.418841, 4.430817, 4.442651, 4.454347, 4.465908,
    4.477337, 4.488636, 4.499810, 4.510859, 4.521789, 4.532599, 4.543295, 4.553877,
    4.564348, 4.574711, 4.584968, 4.595119, 4.605170, 4.615121, 4.624973, 4.634729,
    4.644391, 4.653960, 4.663439, 4.672829, 4.682131, 4.691348, 4.700480, 4.709530,
    4.718499, 4.727388, 4.736198, 4.744932, 4.753591, 4.762174, 4.770685, 4.779124,
    4.787492, 4.795791, 4.804021, 4.812184, 4.820282, 4.828314, 4.836282, 4.844187,
    4.852030,
]


def main():
    out = ["// GENERATED by tools/gen_tables.py -- do not edit.\n"
           "// Twiddle / window / lookup tables computed from their definitions;\n"
           "// validated against the reference's literals at generation time.\n"]
    # ---- Ooura 128 (AEC3): float32 cos/sin as in the original apm_rdft.c.
    cosf = lambda x: f32(math.cos(float(x)))
    sinf = lambda x: f32(math.sin(float(x)))
    # The 128-point tables are published literals (the original run-time
    # initialisation used a libm cosf that was not correctly rounded, so a few
    # entries differ in the last bit from the formula); they are numeric data
    # and are kept verbatim in RDFT_W128 below.
    w128 = [f32(x) for x in RDFT_W128[:32]]
    c128 = [f32(x) for x in RDFT_W128[32:]]
    wf = makewt(32, cosf, sinf) + makect(32, cosf, sinf)
    assert max(abs(float(a) - float(b)) for a, b in zip(wf, w128 + c128)) < 1e-7
    check("rdft_w", w128 + c128, "common_audio/third_party/ooura/fft_size_128/ooura_fft_tables_common.h", "rdft_w")
    tw128 = twiddles(w128, 16)
    # SSE2 tables hold the same numbers in (re,im)-lane layout.
    for nm, idx, sgn in (("rdft_wk1r", 0, 0), ("rdft_wk2r", 2, 0), ("rdft_wk3r", 4, 0),
                         ("rdft_wk1i", 1, 1), ("rdft_wk2i", 3, 1), ("rdft_wk3i", 5, 1)):
        flat = []
        for q in range(16):
            v = tw128[q][idx]
            flat += ([f32(-v), v] if sgn else [v, v])
        check(nm, flat, "common_audio/third_party/ooura/fft_size_128/ooura_fft_tables_neon_sse2.h", nm)
    out.append(farr("kTw128", [x for t in tw128 for x in t]))
    out.append(farr("kRc128", c128, 8))
    p128, _ = bitrv_perm(128)
    out.append(iarr("kBitrv128", p128))
    # ---- fft4g 256 (NS): tables built at run time by the reference with
    # delta in float and cos/sin evaluated in double then narrowed.
    w256 = makewt(64, cosf, sinf)
    c256 = makect(64, cosf, sinf)
    tw256 = twiddles(w256, 32)
    out.append(farr("kW256", w256 + c256, 8))
    out.append(farr("kTw256", [x for t in tw256 for x in t]))
    out.append(farr("kRc256", c256, 8))
    p256, _ = bitrv_perm(256)
    out.append(iarr("kBitrv256", p256))
    # ---- windows
    r8 = lambda x: f32(float("%.8f" % x))
    hann64 = [r8(0.5 * (1 - math.cos(2 * math.pi * k / 63))) for k in range(64)]
    check("kHanning64", hann64, "modules/audio_processing/aec3/aec3_fft.cc", "kHanning64")
    out.append(farr("kHanning64", hann64, 8))
    sq128 = [f32(float("%.14f" % math.sin(math.pi * k / 128))) for k in range(128)]
    check("kSqrtHanning128", sq128, "modules/audio_processing/aec3/aec3_fft.cc", "kSqrtHanning128")
    check("kSqrtHanning", sq128, "modules/audio_processing/aec3/suppression_filter.cc", "kSqrtHanning")
    out.append(farr("kSqrtHanning128", sq128, 8))
    nswin = [r8(math.sin(math.pi * k / 192)) for k in range(96)]
    check("kBlocks160w256FirstHalf", nswin, "modules/audio_processing/ns/noise_suppressor.cc", "kBlocks160w256FirstHalf")
    out.append(farr("kNsWindow96", nswin, 8))
    # NS log(i) table: published 6-decimal literals of unknown rounding
    # provenance (neither round-to-nearest nor truncation of log(i)); numeric
    # data kept verbatim in NS_LOG_TABLE, entries 0..4 are 0 by definition.
    logt = [f32(x) for x in NS_LOG_TABLE]
    logt[10] = f32(math.log(10.0))
    assert all(abs(float(logt[i]) - math.log(i)) < 2e-6 for i in range(5, 129))
    check("log_table", logt, "modules/audio_processing/ns/noise_estimator.cc", "log_table")
    out.append(farr("kNsLogTable", logt, 8))
    s2s = [f32(float("%.7f" % (math.sqrt(2) * math.sin(2 * math.pi * i / 32)))) for i in range(32)]
    s2s[8] = f32(math.sqrt(2.0)); s2s[24] = f32(-f32(math.sqrt(2.0)))
    check("kSqrt2Sin", s2s, "modules/audio_processing/aec3/comfort_noise_generator.cc", "kSqrt2Sin")
    out.append(farr("kSqrt2Sin", s2s, 8))
    # ---- comfort-noise LCG jump-ahead: seed_k = (A_k*seed + C_k) mod 2^31 after k steps
    A, C, M = 69069, 1, 1 << 31
    ak, ck = [1], [0]
    for k in range(1, 64):
        ak.append((ak[-1] * A) % M)
        ck.append((ck[-1] * A + C) % M)
    out.append(iarr("kLcgA", ak, "unsigned", 8))
    out.append(iarr("kLcgC", ck, "unsigned", 8))
    open(OUT, "w").write("\n".join(out))
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
