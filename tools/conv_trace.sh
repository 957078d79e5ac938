#!/bin/bash
# Timeline of one CTA of conv_tc_kernel (clock64 stamps per M tile): builds a separate library with -DMDS_CONV_TRACE here
# (no GPU needed), then on the GPU box: MDS_LIB=tools/bin/libmds_trace.so python tools/bench_conv.py 10 2 2> gpurun_out/trace.txt
mkdir -p tools/bin
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DMDS_CONV_TRACE \
     -o tools/bin/libmds_trace.so ball_action_spotting_b200/csrc/mds_api.cu
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -DMDS_GEMM_TRACE \
     -o tools/bin/libmds_gtrace.so ball_action_spotting_b200/csrc/mds_api.cu
