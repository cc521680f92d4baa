set -x
python -m pytest tests/test_gpu_view_list_split.py -x -q 2>&1 | tail -15
python bench.py --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench2.json')); print(d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['serialized_launch_ms'], d['e2e']['value'], d['e2e']['frac_of_transfer_ceiling'], d['e2e_files']['value'], d['extras']['configs'])"
