mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "conv3x3 or conv_tc" 2>&1 | tail -5
for m in 2 1; do
  timeout 300 python tools/profile_layers.py --batch 4 --conv-mode $m > gpurun_out/layers_conv$m.txt 2>&1
  echo "conv-mode $m"; grep -E "^batch|^stem|^b0|^b1|^b2" gpurun_out/layers_conv$m.txt
done
