# Round-2 measurement pass on one B200 (run through gpurun); raw outputs land in gpurun_out/, summaries are made
# from them by scripts/make_profiles.py r2.  PART=a (tests, bench, configs) | b (launch list + ncu of the cACGMM kernel) | c (ncu of the complex Watson and
# post-fit kernels) | all.  gpurun copies at most 64 MiB back per call: run b and c as separate calls.
set -x
PART=${1:-all}
mkdir -p gpurun_out
if [ "$PART" = a ] || [ "$PART" = all ]; then
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_r2.txt; cat gpurun_out/pytest_gpu_r2.txt
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 900 gpurun_out/bench_r2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_r2.json 2> gpurun_out/bench_reference_r2.err; tail -c 400 gpurun_out/bench_reference_r2.json
(timeout 400 python scripts/time_configs.py; timeout 300 python scripts/run_c3.py; timeout 200 python scripts/c3_parts.py) > gpurun_out/configs_r2.txt 2>&1; tail -12 gpurun_out/configs_r2.txt
fi
if [ "$PART" = b ] || [ "$PART" = all ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_bench_r2.log 2>&1; tail -2 gpurun_out/launches_r2.csv | cut -c 1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_ws_kernel -c 1 -o gpurun_out/em_ws_r2 -f python scripts/one_fit.py 100 > gpurun_out/ncu_ws_r2.log 2>&1; tail -2 gpurun_out/ncu_ws_r2.log
fi
if [ "$PART" = c ] || [ "$PART" = all ]; then
# (reports of these two exceed the 64 MiB gpurun brings back: the text pages are made on the box and the reports dropped)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:em_persistent_kernel -c 1 -o /tmp/cw_r2 -f python scripts/one_fit_cw.py 100 > gpurun_out/ncu_cw_r2.log 2>&1; tail -2 gpurun_out/ncu_cw_r2.log
ncu -i /tmp/cw_r2.ncu-rep --page details > gpurun_out/cw_r2_details.txt 2>&1
ncu -i /tmp/cw_r2.ncu-rep --page raw --csv > gpurun_out/cw_r2_raw.csv 2>&1
python scripts/ncu_breakdown.py /tmp/cw_r2.ncu-rep 25 > gpurun_out/cw_r2_source_breakdown.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'dhtv_cluster_kernel|em_fast_kernel|gev_kernel|apply_bf_kernel' -c 4 -o /tmp/postfit_r2 -f python scripts/run_c3.py --iterations 5 > gpurun_out/ncu_postfit_r2.log 2>&1; tail -2 gpurun_out/ncu_postfit_r2.log
ncu -i /tmp/postfit_r2.ncu-rep --page details > gpurun_out/postfit_r2_details.txt 2>&1
ncu -i /tmp/postfit_r2.ncu-rep --page raw --csv > gpurun_out/postfit_r2_raw.csv 2>&1
ls -la gpurun_out
fi
