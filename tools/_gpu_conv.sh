for r in 12 16 23; do
MDS_DW_ROWS=$r timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows $r b4 value', round(j['value'],1), 'dw2d ms', round(j['roofline_dw']['ms_per_step'],4), 'frac', round(j['roofline_dw']['frac'],3), 'dw3d', round(j['roofline_dw']['dw3d']['ms_per_step'],4))"
done
MDS_DW_ROWS3D=12 timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows3d 12 b4 value', round(j['value'],1), 'dw3d', round(j['roofline_dw']['dw3d']['ms_per_step'],4))"
MDS_DW_ROWS=23 timeout 200 python bench.py --no-cpu-baseline --batch 32 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('rows 23 b32 value', round(j['value'],1), 'dw2d ms', round(j['roofline_dw']['ms_per_step'],4), 'frac', round(j['roofline_dw']['frac'],3))"
