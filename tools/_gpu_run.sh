mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/bench_predictor.py 600 0 2>&1 | tail -1 | tee gpurun_out/r2_predictor.json
timeout 300 python tools/bench_predictor.py 400 1 2>&1 | tail -1 | tee gpurun_out/r2_predictor_tta.json
