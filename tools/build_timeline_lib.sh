#!/usr/bin/env bash
# development: instrumented (SPF_TIMELINE) build of the library as libspurfies_b200_tl.so, next to the normal one.
# Use with SPF_LIBRARY=spurfies_b200/csrc/libspurfies_b200_tl.so python tools/timeline_color.py
set -euo pipefail
cd "$(dirname "$0")/../spurfies_b200/csrc"
A="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
nvcc $A -DSPF_TIMELINE ${SPF_DBGMODE:+-DSPF_DBGMODE} -c mlp_tc2.cu -o /tmp/mlp_tc2_tl.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libspurfies_b200_tl.so grid.o render.o mlp_f32.o mlp_tc.o /tmp/mlp_tc2_tl.o optim.o mesh.o ingest.o local_loss.o
ls -la libspurfies_b200_tl.so
