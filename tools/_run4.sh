set -x
nvidia-smi -L > gpurun_out/r2_n2_gpus.txt
nvidia-smi topo -m > gpurun_out/r2_n2_topo.txt
python -m pytest tests/test_gpu_view_list_split.py -x -q > gpurun_out/r2_pytest_split_n2.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_split_n2.log
tail -15 gpurun_out/r2_pytest_split_n2.log
python tools/time_split.py --devices 1 2 --reps 8 > gpurun_out/r2_split_n2.jsonl 2> gpurun_out/r2_split_n2.err; tail -3 gpurun_out/r2_split_n2.err
cat gpurun_out/r2_split_n2.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/pcie_ceiling.py > gpurun_out/r2_ceiling_n2.jsonl 2>> gpurun_out/r2_ceiling.err
tail -5 gpurun_out/r2_ceiling.err
cat gpurun_out/r2_ceiling_n2.jsonl | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -3 gpurun_out/r2_bench_n2.err
cat gpurun_out/r2_bench_n2.json
