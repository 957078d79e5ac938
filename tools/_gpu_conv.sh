mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/profile_layers.py --batch 4 > gpurun_out/layers_b4.txt 2>&1
grep -E "^batch|^stem|^b0|^b1|^b2|b3.1|b4.1|b5.1|c3d.1|proj" gpurun_out/layers_b4.txt
timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b4 value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['clocks'], {k:(round(v['ms_per_step'],3)) for k,v in j['roofline_by_kind'].items()})"
timeout 300 python bench.py --no-cpu-baseline --batch 32 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b32 value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['clocks'], {k:(round(v['ms_per_step'],3)) for k,v in j['roofline_by_kind'].items()})"
