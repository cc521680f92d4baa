set -x
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu1.log
python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 1 2 --seg-chunks 1 2 4 8 16 --tag minb6 > gpurun_out/r2_sweep1.jsonl 2> gpurun_out/r2_sweep1.err
P2P_B200_LIB=build/variants/libp2p_minb5.so python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 2 4 8 --tag minb5 >> gpurun_out/r2_sweep1.jsonl 2>> gpurun_out/r2_sweep1.err
P2P_B200_LIB=build/variants/libp2p_minb4.so python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 2 4 8 --tag minb4 >> gpurun_out/r2_sweep1.jsonl 2>> gpurun_out/r2_sweep1.err
tail -3 gpurun_out/r2_pytest_gpu1.log
cat gpurun_out/r2_sweep1.jsonl
