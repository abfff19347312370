# round-end evidence run (one B200): tests, bench variants, ncu launch lists and full captures -> gpurun_out/final
# (tools/summarize_final.py turns it into profiles/rNN_final_*)
set -x
mkdir -p gpurun_out/final
(timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/final/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/final/pytest_gpu.log)
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/final/clocks.csv &
SMI=$!
python bench.py > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/final/bench_reference.json 2>&1
X="--no-cpu-baseline --no-config4 --no-pipeline --view-sharded-views 0"
python bench.py --cfg $X > gpurun_out/final/bench_cfg.json 2>&1
python bench.py --cfg --no-batch-cfg $X > gpurun_out/final/bench_cfg_two_forwards.json 2>&1
python bench.py --variant b $X > gpurun_out/final/bench_variant_b.json 2>&1
python bench.py --mv-block standard $X > gpurun_out/final/bench_standard.json 2>&1
python bench.py --scenes-per-gpu 8 --steps 10 $X > gpurun_out/final/bench_8scenes.json 2>&1
kill $SMI
python tools/bench_vae.py > gpurun_out/final/bench_vae.json 2>&1
python tools/scale_check.py > gpurun_out/final/scale_check.txt 2>&1
python tools/prof_gemm.py 8 320 320 32 > gpurun_out/final/prof_ops.txt 2>&1
python tools/prof_gemm.py 64 320 320 32 >> gpurun_out/final/prof_ops.txt 2>&1
python tools/prof_attn.py 1 8192 40 >> gpurun_out/final/prof_ops.txt 2>&1
python tools/prof_attn.py 8 8192 40 3 >> gpurun_out/final/prof_ops.txt 2>&1
python tools/excess.py 8 > gpurun_out/final/excess_v8.txt 2>&1
NCU="--no-graph --no-cpu-baseline --no-config4 --no-pipeline --view-sharded-views 0"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final/launches_cold.csv python bench.py --steps 1 --warmup 3 $NCU > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none --csv --log-file gpurun_out/final/launches_warm.csv python bench.py --steps 1 --warmup 3 $NCU > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/final/ncu_gemm_conv_l0 python tools/prof_gemm.py 8 320 320 32 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn64q -s 3 -c 1 -o gpurun_out/final/ncu_attn_l0 python tools/prof_attn.py 1 8192 40 5 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_flat -s 3 -c 1 -o gpurun_out/final/ncu_gn_flat_l0 python tools/prof_gn.py 8 1024 320 0 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.log 2>&1
tail -n 3 gpurun_out/final/pytest_gpu.log gpurun_out/final/smoke.log
