for v in NOCOORD NOSTORE NOBLEND NOCOORD_NOBLEND NOCOORD_NOSTORE TEXONLY; do
P2P_B200_LIB=build/variants/libp2p_$v.so python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 2 --tag $v >> gpurun_out/r2_ablation.jsonl 2>> gpurun_out/r2_ablation.err
done
python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 2 --tag full >> gpurun_out/r2_ablation.jsonl 2>> gpurun_out/r2_ablation.err
cat gpurun_out/r2_ablation.jsonl
