mkdir -p gpurun_out
for m in 3 0; do
for B in 4 32; do
timeout 300 python bench.py --no-cpu-baseline --tail-mode $m --batch $B > gpurun_out/r2_bench_m${m}_b$B.json 2> gpurun_out/r2_bench_m${m}_b$B.err
done; done
python - <<'PY'
import json
for f in ("r2_bench_m3_b4","r2_bench_m0_b4","r2_bench_m3_b32","r2_bench_m0_b32"):
    try:
        j=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", round(j["value"],1), "e2e", round(j["e2e"]["value"],1), "launches", j["gpu_launches"], {k: round(v["ms_per_step"],3) for k,v in j["roofline_by_kind"].items()})
        print("   dw", round(j["roofline_dw"]["achieved"]), round(j["roofline_dw"]["frac"],3), "3d", round(j["roofline_dw"]["dw3d"]["achieved"]), round(j["roofline_dw"]["dw3d"]["frac"],3))
    except Exception as e:
        print(f, "ERR", e, open(f"gpurun_out/{f}.err").read()[-1500:])
PY
for cfg in "12 8" "23 12" "46 23"; do
set -- $cfg
echo "== mode 3 MDS_DW_ROWS=$1 MDS_DW_ROWS3D=$2"
MDS_DW_ROWS=$1 MDS_DW_ROWS3D=$2 timeout 300 python tools/profile_layers.py --batch 4 --tail-mode 3 2>&1 | grep -E "batch|b3.1.dw|b4.1.dw|b5.0.dw|b5.1.dw|c3d.1.dw"
done
