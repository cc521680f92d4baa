set -x
nvidia-smi -L > gpurun_out/r2_n8_gpus.txt
nvidia-smi topo -m > gpurun_out/r2_n8_topo.txt
lscpu | grep -i 'numa\|model name\|socket\|^CPU(s)' > gpurun_out/r2_n8_lscpu.txt
free -g >> gpurun_out/r2_n8_lscpu.txt
python -m pytest tests/test_gpu_view_list_split.py -x -q > gpurun_out/r2_pytest_split_n8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_split_n8.log
tail -5 gpurun_out/r2_pytest_split_n8.log
for n in 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/pcie_ceiling.py --steps 2 > gpurun_out/r2_ceiling_n$n.jsonl 2>> gpurun_out/r2_ceiling.err
cat gpurun_out/r2_ceiling_n$n.jsonl | grep -v topology | cut -c1-300
done
python tools/time_split.py --devices 1 2 4 8 --reps 8 > gpurun_out/r2_split_n8.jsonl 2> gpurun_out/r2_split_n8.err; tail -3 gpurun_out/r2_split_n8.err
cat gpurun_out/r2_split_n8.jsonl | cut -c1-700
python tools/time_folder.py --files 32 --devices 1 2 4 8 --formats jpg --repeat 2 > gpurun_out/r2_folder_devices.jsonl 2> gpurun_out/r2_folder_devices.err; tail -3 gpurun_out/r2_folder_devices.err
cat gpurun_out/r2_folder_devices.jsonl
for n in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; tail -3 gpurun_out/r2_bench_n$n.err
cat gpurun_out/r2_bench_n$n.json | cut -c1-5000
done
