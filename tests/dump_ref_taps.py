"""Helper process for test_stage_taps: runs the dump variant of the compiled reference
(oracle/_ref/libwap_ref_dump.so, -DWEBRTC_APM_DEBUG_DUMP=1) on one synthetic leg so that every
ApmDataDumper::DumpRaw tap lands in <outdir>/<name>_<instance>-<reinit>.dat.
usage: dump_ref_taps.py <outdir> <frames> <leg>"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from common import synthetic_leg  # noqa: E402

outdir, nf, leg = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
L = C.CDLL(os.path.join(HERE, "..", "oracle", "_ref", "libwap_ref_dump.so"))
L.ref_apm_create.restype = C.c_void_p
L.ref_apm_create.argtypes = [C.c_int] * 7
L.ref_apm_run_i16.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_void_p]
L.ref_dump_activate.argtypes = [C.c_char_p]
assert L.ref_dump_activate(outdir.encode()) == 1
far, near = synthetic_leg(leg, nf)
h = L.ref_apm_create(1, 1, 1, 32000, 0, 0, 0)
out = np.zeros(nf * 160, np.int16)
err = L.ref_apm_run_i16(h, 16000, 1, 1, nf, far.ctypes.data_as(C.c_void_p), near.ctypes.data_as(C.c_void_p),
                        out.ctypes.data_as(C.c_void_p), 0, None)
sys.exit(err)
