set -x
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
python tools/dbg/train_timeline.py
TB=2048 python tools/dbg/train_timeline.py
python bench.py > gpurun_out/bench_r1_final_1gpu.json 2> gpurun_out/bench_r1_final_1gpu.err
tail -c 400 gpurun_out/bench_r1_final_1gpu.json
