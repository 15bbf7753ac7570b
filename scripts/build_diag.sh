#!/bin/bash
# Side libraries for scripts/diag_ssb.py: libadvchain_b200.so with the squaring-step adjoint taken apart
# (ADVK_SSB_DIAG = 1..4, see advk_morph.cu).  WRONG results by construction -- timing only; they live
# under scripts/_diag/ (git-ignored, shipped to the GPU box by gpurun) and are never loaded by the package.
set -e
cd "$(dirname "$0")/.."
python -m advchain_b200.build > /dev/null
mkdir -p scripts/_diag
OBJS=$(ls advchain_b200/csrc/_obj/*.o | grep -v advk_morph.o)
for k in 1 2 3 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DADVK_SSB_DIAG=$k \
      -c advchain_b200/csrc/advk_morph.cu -o scripts/_diag/advk_morph_diag$k.o &
done
wait
for k in 1 2 3 4; do
  nvcc -shared -o scripts/_diag/libadvk_diag$k.so $OBJS scripts/_diag/advk_morph_diag$k.o -lcudart
done
ls -la scripts/_diag/*.so
