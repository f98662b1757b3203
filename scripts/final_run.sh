set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_r1.txt; cat gpurun_out/pytest_gpu_r1.txt
timeout 600 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; tail -c 600 gpurun_out/bench_r1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_r1.json 2> gpurun_out/bench_reference_r1.err; tail -c 400 gpurun_out/bench_reference_r1.json
(timeout 300 python scripts/time_configs.py; timeout 300 python scripts/run_c3.py) > gpurun_out/configs_r1.txt 2>&1; tail -8 gpurun_out/configs_r1.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_bench.log 2>&1; tail -2 gpurun_out/launches_r1.csv | cut -c 1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_ws_kernel -c 1 -o gpurun_out/em_ws_r1 -f python scripts/one_fit.py 100 > gpurun_out/ncu_ws.log 2>&1; tail -2 gpurun_out/ncu_ws.log
