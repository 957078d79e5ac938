#!/bin/bash
# One `ncu --set full` launch of every kernel kind of the FULL forward at batch B (default 4 = BASELINE.json configs[1]).
# Run on the GPU box (gpurun); reports land in gpurun_out/ncu_r02_*.ncu-rep and are summarised by tools/ncu_summarize.py.
B=${1:-4}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
KEEP=${KEEP_REPS:-"gemm_b41 conv_tc"}     # gpurun brings back at most 64 MiB: keep the raw / source CSV pages of every capture, the .ncu-rep of a few
cap() {  # name, kernel regex, skip, count
  $NCU -k "regex:$2" -s $3 -c $4 -o $OUT/ncu_r02_$1 python tools/ncu_target.py $B > $OUT/ncu_r02_$1.log 2>&1 || tail -3 $OUT/ncu_r02_$1.log
  ncu -i $OUT/ncu_r02_$1.ncu-rep --page raw --csv > $OUT/ncu_r02_$1.raw.csv 2>/dev/null
  case " $KEEP " in *" $1 "*) ;; *) rm -f $OUT/ncu_r02_$1.ncu-rep ;; esac
}
cap stem 'stem_tc_kernel' 0 1            # TMA + tcgen05 stem (first half of the images: the encoder runs as two halves on two streams)
cap conv_tc 'conv_tc_kernel' 0 4         # blocks.0.0 (column taps folded into N), 1.0, 1.1, 2.0
cap conv_tc_ws 'conv_tc_ws_kernel' 0 1   # blocks.2.1 (3x3 weights streamed from L2)
cap gemm_b41 'gemm_tc_kernel' 8 2         # blocks.4.1 conv_pw (resident A), conv_pwl (streamed, pre-gated weights)
cap gemm_b51 'gemm_tc_kernel' 18 2        # blocks.5.1 conv_pw, conv_pwl
cap dw_b30 'dwconv_tma_kernel' 0 1        # blocks.3.0 stride 2
cap dw_b41 'dwconv_tma_kernel' 4 1        # blocks.4.1
cap dw_b51 'dwconv_tma_kernel' 9 1        # blocks.5.1
cap dw_3d 'dwconv_tma_kernel<\(int\)3' 0 1
cap se_b41 'se_fc_kernel' 4 1
cap se_b51 'se_fc_kernel' 9 1
cap head 'gem_kernel|gem_finish_kernel|linear_head_kernel' 0 3
# launch list of the bench command (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/ncu_r02_launches_b$B.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch $B > $OUT/ncu_r02_launches_b$B.log 2>&1
ls -la $OUT/ncu_r02_* | awk '{print $5, $9}'
