set -x
ncu --set full --clock-control none --import-source on -k regex:project_rows -s 10 -c 1 -f -o gpurun_out/r2_prof_rows_seg2 python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 2 --batch 8 --steps 2 > gpurun_out/r2_ncu_rows.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_rows -s 10 -c 1 -f -o gpurun_out/r2_prof_rows_seg1 python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 1 --batch 8 --steps 2 >> gpurun_out/r2_ncu_rows.log 2>&1
tail -5 gpurun_out/r2_ncu_rows.log
