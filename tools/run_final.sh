set -x
python bench.py > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.log 2>&1
python bench.py --cfg --no-cpu-baseline > gpurun_out/bench_cfg.log 2>&1
python bench.py --scenes-per-gpu 8 --steps 10 --no-cpu-baseline > gpurun_out/bench_s8.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -n 2 gpurun_out/smoke.log
