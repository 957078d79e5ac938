mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --mode sweep --gpus $N > gpurun_out/r2_sweep_full_n$N.json 2> gpurun_out/r2_sweep_full_n$N.err
  tail -c 300 gpurun_out/r2_sweep_full_n$N.err | grep -v "^$" | tail -2
done
timeout 600 $TR --nproc-per-node 8 --master-port 29611 bench.py --mode sweep --gpus 8 --tta 1 > gpurun_out/r2_sweep_full_tta_n8.json 2> gpurun_out/r2_sweep_full_tta_n8.err
for N in 8 4 2; do
  timeout 300 $TR --nproc-per-node $N --master-port $((29700+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_b4_n$N.json 2> gpurun_out/r2_bench_b4_n$N.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2_sweep_full*_n[248].json"))+sorted(glob.glob("gpurun_out/r2_bench_b4_n[248].json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], "n", j["n_gpus"], "value", round(j["value"],1), "ms/step", round(j["ms_per_step"],2), "e2e", (round(j["e2e"]["value"],1), round(j["e2e"].get("h2d_ceiling_stacks_per_s",0),1), round(j["e2e"].get("h2d_ceiling_gbs",0),1)) if j.get("e2e") else None, j.get("sharded_equals_unsharded",{}).get("bit_identical"), j.get("spots"))
    except Exception as e:
        print(f, "ERR", e)
PY
