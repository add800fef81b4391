mkdir -p gpurun_out/r02d
cd /root/repo
O=gpurun_out/r02d
(timeout 900 python -m pytest tests/test_gpu_zy_sk_gemm.py -x -q -m gpu -k "ns_gemm" > $O/pytest_ns.log 2>&1; echo "exit $?" >> $O/pytest_ns.log)
grep -v "^  File\|^Extension\|^$" $O/pytest_ns.log | tail -8
(timeout 300 python tools/dep_trace.py --B 104 > $O/dep_trace_v2_b104.json 2> $O/dep_trace_v2_b104.err; echo "exit $?" >> $O/dep_trace_v2_b104.err)
(B200_DEP_FUSED=2 timeout 300 python tools/dep_trace.py --B 1 > $O/dep_trace_v2_b1.json 2> $O/dep_trace_v2_b1.err)
python -c "
import json
for f in ['$O/dep_trace_v2_b104.json','$O/dep_trace_v2_b1.json']:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['total_us']); print(json.dumps(d.get('detail'),indent=1))
"
tail -3 $O/dep_trace_v2_b104.err
(timeout 600 python tools/kbench.py --what gemm --M 104 --ns-sweep --only temporal > $O/kbench_ns.jsonl 2> $O/kbench_ns.err; echo "exit $?" >> $O/kbench_ns.err)
(timeout 300 python tools/kbench.py --what gemm --M 104 --ns-sweep --only depformer_in_all >> $O/kbench_ns.jsonl 2>> $O/kbench_ns.err)
(timeout 300 python -m pytest tests/test_gpu_lm.py -x -q -m gpu -k "shortened or fused or greedy" > $O/pytest_lm.log 2>&1; echo "exit $?" >> $O/pytest_lm.log)
grep -v "^  File\|^Extension\|^$" $O/pytest_lm.log | tail -8
timeout 500 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['lm_step']['ms'],d['roofline_gemm']['frac'], d['kv_fill_sweep'], d['secondary'])"
tail -5 $O/bench.err
