set -x
for st in 1 2 3; do
python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 1 2 --seg-chunks 2 4 --streams $st --tag streams >> gpurun_out/r2_sweep3.jsonl 2>> gpurun_out/r2_sweep3.err
done
cat gpurun_out/r2_sweep3.jsonl | cut -c1-250
python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench1.err
cat gpurun_out/r2_bench1.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench1_ref.json 2> gpurun_out/r2_bench1_ref.err; cat gpurun_out/r2_bench1_ref.json | cut -c1-300
