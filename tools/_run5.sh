set -x
python -m pytest tests/test_gpu_view_list_split.py -x -q > gpurun_out/r2_pytest_split_n1.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_split_n1.log
tail -25 gpurun_out/r2_pytest_split_n1.log
python tools/pcie_ceiling.py > gpurun_out/r2_ceiling_n1.jsonl 2> gpurun_out/r2_ceiling.err
tail -5 gpurun_out/r2_ceiling.err
cat gpurun_out/r2_ceiling_n1.jsonl | cut -c1-600
python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 1 2 --seg-chunks 2 3 4 5 --tag segs > gpurun_out/r2_sweep2.jsonl 2> gpurun_out/r2_sweep2.err
cat gpurun_out/r2_sweep2.jsonl | cut -c1-220
python tools/time_split.py --devices 1 --reps 8 > gpurun_out/r2_split_n1.jsonl 2> gpurun_out/r2_split_n1.err; tail -3 gpurun_out/r2_split_n1.err
cat gpurun_out/r2_split_n1.jsonl
