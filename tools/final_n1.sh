#!/bin/bash
# The round's single-GPU measurement session: parity log, bench lines, per-scene lines and sweeps, launch list, ncu captures.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -k "per_path or film_matches or full_frame or deep_paths or kernel_selection" 2>&1 | grep -v "^$" | tail -90 > gpurun_out/r2_parity_gpu_tests.txt
python -m pytest tests -m gpu -q 2>&1 | tail -3 >> gpurun_out/r2_parity_gpu_tests.txt
tools/run_scaling.sh r2_n1 1 --steps 5 --warmup 3 --spp-sweep 1,4,16,64,256
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_reference_arm.err
tools/run_scaling.sh r2_n1_cornell_plane_light 1 --scene cornell_plane_light --spp 1024 --steps 5 --warmup 3 --spp-sweep 1,4,16,64,256,1024
tools/run_scaling.sh r2_n1_cornell_large_box_256spp 1 --scene cornell_large_box --spp 256 --steps 5 --warmup 3 --no-cpu-baseline
tools/run_scaling.sh r2_n1_stress_all_256spp 1 --scene stress_all --spp 256 --steps 3 --warmup 3 --no-cpu-baseline
tools/run_scaling.sh r2_n1_classed_all_256spp 1 --scene classed_all --spp 256 --steps 5 --warmup 3 --no-cpu-baseline
# launch list of a bench run (every launch with its device time; shares, not absolutes)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
# the three render kernels, one launch each (33.5 M camera paths)
for job in "init_cornell r2_fast_init_cornell" "cornell_plane_light r2_classed_plane_light" "stress_all r2_general_stress_all"; do
  set -- $job
  python tools/profile_render.py $1 1024 1024 32 > gpurun_out/plain_$2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/$2 python tools/profile_render.py $1 1024 1024 32 > gpurun_out/ncu_$2.log 2>&1
  cat gpurun_out/plain_$2.log
done
