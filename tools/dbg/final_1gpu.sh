set -x
python bench.py > gpurun_out/bench_r1_final_1gpu.json 2> gpurun_out/bench_r1_final_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_launch4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_ -c 3 -o gpurun_out/prof_r1d python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full4.log 2>&1
ncu -i gpurun_out/prof_r1d.ncu-rep --page raw --csv > gpurun_out/prof_r1d_raw.csv 2>/dev/null
tail -c 600 gpurun_out/bench_r1_final_1gpu.json
