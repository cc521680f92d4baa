set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu2.log
tail -4 gpurun_out/r2_pytest_gpu2.log
# racecheck: projection kernels (shuffles, no shared memory), JPEG / PNG encoders and the JPEG decoder (shared-memory scans, atomicOr emitters)
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest -q -x \
  "tests/test_gpu_view_list_split.py::test_cube_faces_one_launch_bit_exact_against_oracle" \
  "tests/test_gpu_view_list_split.py::test_row_bands_tile_the_view" \
  "tests/test_gpu_jpeg.py::test_encode_batch_and_extremes" \
  "tests/test_gpu_jpeg.py::test_project_views_jpeg_equals_imwrite_of_the_views" \
  "tests/test_gpu_jpeg.py::test_device_huffman_stage_is_used_and_equals_host_stage" \
  "tests/test_gpu_png.py::test_encode_png_batch_with_noise" \
  "tests/test_gpu_png.py::test_small_images_stored_blocks_and_chunk_boundaries" \
  > gpurun_out/r2_compute_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_compute_sanitizer_racecheck.txt
tail -12 gpurun_out/r2_compute_sanitizer_racecheck.txt
# ncu: launch list of the bench command, then one full capture of the shipped kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:project_rows -s 10 -c 1 -f -o gpurun_out/r2_prof_rows_seg4 python tools/sweep_variants.py --samplers 1 --warp-ws 32 --nbs 1 --mirrors 2 --seg-chunks 4 --batch 8 --steps 2 > gpurun_out/r2_ncu_rows4.log 2>&1
ls -la gpurun_out/r2_prof_rows_seg4.ncu-rep
