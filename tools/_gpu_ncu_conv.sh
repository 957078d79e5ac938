OUT=gpurun_out
mkdir -p $OUT
ncu --set full --clock-control none --import-source on -f -k "regex:conv_tc_kernel" -c 4 -o $OUT/ncu_r02_conv_tc python tools/ncu_target.py 4 > $OUT/ncu_r02_conv_tc.log 2>&1 || tail -3 $OUT/ncu_r02_conv_tc.log
ncu -i $OUT/ncu_r02_conv_tc.ncu-rep --page raw --csv > $OUT/ncu_r02_conv_tc.raw.csv 2>/dev/null
ls -la $OUT/ncu_r02_conv_tc*
