#!/bin/bash
# files flow on all GPUs of the box at once, one process per GPU (what the bench's e2e_files does under torchrun), swept
# over images in flight / decoder mode / host wait mode.  usage: tools/_run8.sh <n_gpus>
N=${1:-8}
mkdir -p gpurun_out
T0=$(python -c "import time; print(time.time() + 30)")
for i in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$i python tools/time_files_flow.py --threads 4 8 --huffman 3 1 --wait 0 1 --images 96 --t0 $T0 --slot-seconds 2.5 \
    > gpurun_out/ff${N}_$i.jsonl 2> gpurun_out/ff${N}_$i.err &
done
wait
python - <<PY
import json,glob,collections
agg=collections.defaultdict(list)
for f in sorted(glob.glob("gpurun_out/ff${N}_*.jsonl")):
    for l in open(f):
        d=json.loads(l); agg[(d["host_wait"],d["gpu_huffman"],d["threads"])].append((d["ms_per_image"],d["cpu_ms_per_image"]))
for k,v in sorted(agg.items()):
    print("wait %d huffman %d threads %d:" % k, "n=%d ms/img mean %.3f max %.3f cpu %.2f" % (len(v), sum(x for x,_ in v)/len(v), max(x for x,_ in v), sum(c for _,c in v)/len(v)))
PY
