#!/usr/bin/env bash
# Build A/B variants of libdmpnn_sm100.so that differ only in the fused depth step's build-time switches
# (csrc/step_fused_kernel.cuh: DMPNN_MMA_ORDER, DMPNN_H0_DIRECT, DMPNN_WSTAGES); each lands in chemprop_b200/lib/variants/<tag>/libdmpnn_sm100.so.
# tests/native/fused_step_harness picks one up through LD_LIBRARY_PATH (tools/gpu_call.sh step `variants`).
set -eu
cd "$(dirname "$0")/.."
python -c "from chemprop_b200 import build; build.build()"
OBJ=chemprop_b200/lib/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
build_one() {   # tag, -D flags...
  tag=$1; shift
  d=chemprop_b200/lib/variants/$tag
  mkdir -p "$d"
  for f in step_fused step_fused_fwd step_fused_bwd step_fused_far_fwd step_fused_far_bwd step_fused_atom_fwd step_fused_atom_bwd; do
    nvcc $FLAGS "$@" -c chemprop_b200/csrc/$f.cu -o "$d/$f.o" &
  done
  wait
  objs=$(ls $OBJ/*.o | grep -v step_fused)
  nvcc -shared -o "$d/libdmpnn_sm100.so" $objs "$d"/step_fused*.o -gencode arch=compute_100a,code=sm_100a
  rm -f "$d"/step_fused*.o
  echo "built $d"
}
build_one exp  -DDMPNN_EXPERIMENTS=1
if [ "${DMPNN_ALL_VARIANTS:-0}" = 1 ]; then     # round-2 A/Bs that lost (profiles/r2_fused_step_variants.log)
  build_one h0d_w2  -DDMPNN_H0_DIRECT=1 -DDMPNN_WSTAGES=2
  build_one h0d_w3  -DDMPNN_H0_DIRECT=1 -DDMPNN_WSTAGES=3
  build_one h0d_w4  -DDMPNN_H0_DIRECT=1 -DDMPNN_WSTAGES=4
fi
make -s -C tests/native
