python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu5.log
tail -4 gpurun_out/r2_pytest_gpu5.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_rows -s 10 -c 1 -f -o gpurun_out/r2_prof_rows_final python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 4 --batch 8 --steps 2 > gpurun_out/r2_ncu_rows_final.log 2>&1
ls -la gpurun_out/r2_prof_rows_final.ncu-rep
python bench.py > gpurun_out/r2_bench_final_n1.json 2> gpurun_out/r2_bench_final_n1.err; tail -2 gpurun_out/r2_bench_final_n1.err
python - <<'PY'
import json; d = json.load(open('gpurun_out/r2_bench_final_n1.json'))
print('value', d['value'], 'frac', d['roofline']['frac'], 'serial', d['roofline']['serialized_launch_ms'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_transfer_ceiling'], 'files', d['e2e_files']['value'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
print(d['extras']['configs']); print({k: (v['gpu_ms_per_image'], v['byte_identical_to_cpu_flow']) for k, v in d['extras'].items() if k.endswith('files')})
PY
