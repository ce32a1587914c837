set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
python tools/time_sweeps.py > gpurun_out/sweeps.log 2>&1
python tools/time_sweeps.py sigma 3.06 >> gpurun_out/sweeps.log 2>&1
python tools/time_sweeps.py hard >> gpurun_out/sweeps.log 2>&1
python tools/time_sweeps.py highres 16 >> gpurun_out/sweeps.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extra --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sweep_tc|colsum|cand_eval|count_emit|fine_match" -s 7 -c 7 -o gpurun_out/r2_step -f python tools/profile_step.py 64 3 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/r2_step.ncu-rep --page raw --csv > gpurun_out/r2_step_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/sweeps.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
