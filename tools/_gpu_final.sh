mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/gputests_r02.txt; cat gpurun_out/gputests_r02.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.txt 2>&1; tail -2 gpurun_out/smoke_r02.txt
timeout 300 python bench.py > gpurun_out/bench_r02_b4.json 2> gpurun_out/bench_r02_b4.err
timeout 300 python bench.py --batch 32 --no-cpu-baseline > gpurun_out/bench_r02_b32.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02_reference.json 2>/dev/null

timeout 300 python bench.py --mode sweep > gpurun_out/sweep_r02_full_n1.json 2> gpurun_out/sweep_r02_full_n1.err
for f in bench_r02_b4 bench_r02_b32 bench_r02_reference bench_r02_b4_convmode0 sweep_r02_full_n1; do python -c "
import json,sys
j=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(j['value'],2), j.get('unit'), 'e2e', (j.get('e2e') or {}).get('value'), j.get('clocks',{}).get('reasons'), (j.get('roofline') or {}).get('frac'), (j.get('roofline_dw') or {}).get('frac'))"; done
timeout 200 python tools/profile_layers.py --batch 4 > gpurun_out/layers_r02_b4.txt 2>&1
timeout 200 python tools/profile_layers.py --batch 32 > gpurun_out/layers_r02_b32.txt 2>&1
head -1 gpurun_out/layers_r02_b4.txt gpurun_out/layers_r02_b32.txt
timeout 200 python tools/bench_predictor.py 600 0 > gpurun_out/predictor_r02.json 2>/dev/null; tail -1 gpurun_out/predictor_r02.json | cut -c1-300
timeout 900 bash tools/ncu_capture_all.sh 4 2>&1 | tail -20
