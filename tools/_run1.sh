mkdir -p gpurun_out/r02a
cd /root/repo
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a/pytest.log)
timeout 400 python bench.py > gpurun_out/r02a/bench_default.json 2> gpurun_out/r02a/bench_default.err
timeout 300 python tools/kbench.py --what gemm,mimi --M 1,48,104 --B 1,104 > gpurun_out/r02a/kbench.jsonl 2> gpurun_out/r02a/kbench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a/mimi_launches_b104.csv python tools/mimi_frames.py --B 104 --frames 3 > gpurun_out/r02a/mimi_l104.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a/mimi_launches_b1.csv python tools/mimi_frames.py --B 1 --frames 3 > gpurun_out/r02a/mimi_l1.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02a/bench_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-secondary > gpurun_out/r02a/bench_ncu.log 2>&1
tail -3 gpurun_out/r02a/pytest.log; cat gpurun_out/r02a/bench_default.json | head -c 3000
