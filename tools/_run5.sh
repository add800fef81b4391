mkdir -p gpurun_out/r02e
cd /root/repo
O=gpurun_out/r02e
(timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log)
grep -v "^  File\|^Extension\|^$" $O/pytest.log | tail -6
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
python -c "
import json;d=json.loads(open('$O/bench_default.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['e2e']['ms_per_step'],d['lm_step']['ms'],d['mimi'],d['roofline_gemm']['frac'], d['lm_b1']['fill_200']['p50_ms'], d['kv_fill_sweep'])"
(timeout 300 python tools/dep_trace.py --B 104 > $O/dep_trace_b104.json 2> $O/dep_trace_b104.err)
(timeout 300 python tools/dep_trace.py --B 8 > $O/dep_trace_b8.json 2>> $O/dep_trace_b104.err)
python -c "
import json
for f in ['$O/dep_trace_b104.json','$O/dep_trace_b8.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['total_us'], {k:v['avg_us'] for k,v in d['phases'].items()})
"
timeout 300 python tools/kbench.py --what gemm,mimi --M 48,104 --B 1,104 > $O/kbench.jsonl 2> $O/kbench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/mimi_launches_b104.csv python tools/mimi_frames.py --B 104 --frames 3 > $O/mimi_l104.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/bench_launches.csv python bench.py --steps 2 --warmup 1 --skip-cpu-baseline --skip-secondary > $O/bench_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_ns -c 6 -o $O/gemm_ns_full python tools/kbench.py --what gemm --M 104 --only temporal.linear_in --no-legacy > $O/ncu_ns.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mimi_tc_kernel -s 200 -c 12 -o $O/mimi_tc_full python tools/mimi_frames.py --B 104 --frames 4 > $O/ncu_mimi.log 2>&1
ls -la $O
