#!/bin/bash
# Offline (no GPU) look at the tensor-core pass B after a source change: registers / spills of k_hmma<CLAHE>, static
# instruction count and the patterns this round's profile flagged (ptxas re-deriving loop invariants inside the loop).
# usage: tools/sass_stats.sh [tag] [extra nvcc flags, e.g. -DSARPRO_HMMA_PACKED_LUT]   (writes /tmp/hmma_<tag>.{o,sass})
set -e
tag=${1:-cur}; shift || true
here=$(cd "$(dirname "$0")/.." && pwd)
fn='_ZN6sarpro6k_hmmaILb1EEEvNS_11HResizeArgsENS_10HMmaParamsE'
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xptxas -v "$@" -x cu -c "$here/sarpro_b200/csrc/kernels_hmma.cu" -o /tmp/hmma_$tag.o 2> /tmp/hmma_$tag.ptxas
grep -A2 "k_hmmaILb1" /tmp/hmma_$tag.ptxas | grep -i "registers\|spill"
cuobjdump -sass -fun "$fn" /tmp/hmma_$tag.o > /tmp/hmma_$tag.sass
s=/tmp/hmma_$tag.sass
echo "static instructions: $(grep -c '^\s*/\*[0-9a-f]\{4\}\*/' $s)"
echo "IMAD x 0x10001 (packed table base re-derived): $(grep -c '0x10001' $s)   MOV of the clamp bias: $(grep -c 'MOV.*-0x1ff0200' $s)"
echo "FFMA $(grep -c FFMA $s)  LDS.128 $(grep -c 'LDS.128' $s)  LDS.U16 $(grep -c 'LDS.U16' $s)  LDS(32) $(grep -c ' LDS R' $s)  IMMA $(grep -c IMMA $s)  LDL/STL $(grep -c 'LDL\|STL' $s)"
echo "top opcodes:"; grep -o "^\s*/\*[0-9a-f]*\*/\s*\(@!\?U\?P[0-9]\s*\)\?[A-Z0-9_.]*" $s | awk '{print $NF}' | sort | uniq -c | sort -k1 -n -r | head -12 | tr '\n' ' '; echo
