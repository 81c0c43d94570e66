#!/bin/bash
# Float divisions that nvcc turns into reciprocal multiplications under -ftz=true (see fdiv in
# csrc/wap_dev.cuh): PTX of the kernels with and without -ftz; the sites are the div.rn with an
# immediate divisor that only the -ftz=false build has.  Expected output: "0 sites".
cd "$(dirname "$0")/../webrtc-audio-processing_b200" || exit 1
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -I csrc -I ../include"
T=$(mktemp -d)
nvcc $F -ftz=false -ptx csrc/wap_engine.cu -o $T/noftz.ptx || exit 1
grep -B12 "div.rn.f32.*, 0f" $T/noftz.ptx | grep "\.loc\|div.rn" | awk '/\.loc/{l=$0} /div.rn/{print l; print $0}'
echo "$(grep -c 'div.rn.f32.*, 0f' $T/noftz.ptx) sites"
rm -rf $T
