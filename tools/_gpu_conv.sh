mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --batch 32 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b32 value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['clocks'], {k:(round(v['ms_per_step'],3)) for k,v in j['roofline_by_kind'].items()})"
grep sweep16 gpurun_out/parity_report.jsonl | tail -5 | cut -c1-300
