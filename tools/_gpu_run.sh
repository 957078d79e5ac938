mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_r02_b4.json 2> gpurun_out/bench_r02_b4.err
timeout 300 python bench.py --no-cpu-baseline --batch 32 > gpurun_out/bench_r02_b32.json 2> gpurun_out/bench_r02_b32.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
for m in 0 1 2; do timeout 300 python bench.py --no-cpu-baseline --tail-mode $m > gpurun_out/bench_r02_b4_tailmode$m.json 2>/dev/null; done
timeout 300 python tools/profile_layers.py --batch 4 > gpurun_out/layers_r02_b4.txt 2>&1
timeout 300 python tools/profile_layers.py --batch 32 > gpurun_out/layers_r02_b32.txt 2>&1
timeout 600 python bench.py --mode sweep > gpurun_out/sweep_r02_full_n1.json 2> gpurun_out/sweep_r02_full_n1.err
timeout 900 bash tools/ncu_capture_all.sh 4 2>&1 | tail -3
du -sh gpurun_out
