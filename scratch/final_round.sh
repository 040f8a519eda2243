set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench_r01_s8.json 2> gpurun_out/bench_r01_s8.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r01_s8_ref.json 2> gpurun_out/bench_r01_s8_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_r01_v8.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_v8.log 2>&1; echo ncu rc=$?
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python profiles/bench_kernels.py > gpurun_out/kernels_r01_s8.json 2> gpurun_out/kernels_r01_s8.err; echo kernels rc=$?
