mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/gputests_r02.txt; cat gpurun_out/gputests_r02.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('b4 value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['clocks']['reasons'])"
grep sweep16 gpurun_out/parity_report.jsonl | tail -5 | head -2 | cut -c1-120
