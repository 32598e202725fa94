set -x
for cfg in "tiny.en 8" "base.en 1" "base.en 8" "small 8" "medium 1" "large 1" "large 8"; do
  set -- $cfg
  timeout 300 python bench.py --model $1 --batch $2 --steps 5 --warmup 3 --no-cpu-baseline >> gpurun_out/bench_r1e_sweep.jsonl 2>> gpurun_out/bench_r1e_sweep.err || echo "FAILED $cfg" >> gpurun_out/bench_r1e_sweep.jsonl
done
