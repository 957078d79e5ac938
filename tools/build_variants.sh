#!/bin/bash
# libmds_b200.so with every MMA issue order (MDS_NUMERICS_VARIANT, csrc/common.cuh) for tests/parity_variants.py
mkdir -p tools/bin
for v in ${VARIANTS:-0 1 2 3 4 5 6 7}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DMDS_NUMERICS_VARIANT=$v \
       -o tools/bin/libmds_v$v.so ball_action_spotting_b200/csrc/mds_api.cu &
done
wait
ls -la tools/bin/libmds_v*.so | wc -l
