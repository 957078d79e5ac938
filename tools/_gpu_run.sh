mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mbconv_tail_gpu.py -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_e2e_gpu.py -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_b4.json 2> gpurun_out/r2_bench_b4.err
timeout 300 python bench.py --no-cpu-baseline --batch 32 > gpurun_out/r2_bench_b32.json 2> gpurun_out/r2_bench_b32.err
python - <<'PY'
import json
for f in ("r2_bench_b4","r2_bench_b32"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "launches", j["gpu_launches"], {k: round(v["ms_per_step"],3) for k,v in j["roofline_by_kind"].items()})
        print("   dw", round(j["roofline_dw"]["achieved"]), round(j["roofline_dw"]["frac"],3), "3d", round(j["roofline_dw"]["dw3d"]["achieved"]), round(j["roofline_dw"]["dw3d"]["frac"],3))
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
