python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 1 2 --seg-chunks 4 --batch 1 --steps 64 --tag l2warm_b1 2>/dev/null | cut -c1-260
python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 4 --batch 2 --steps 32 --streams 2 --tag l2warm_b2_s2 2>/dev/null | cut -c1-260
python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 4 --batch 16 --steps 8 --streams 2 --tag cold_b16_s2 2>/dev/null | cut -c1-260
