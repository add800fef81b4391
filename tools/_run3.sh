mkdir -p gpurun_out/r02c
cd /root/repo
O=gpurun_out/r02c
(timeout 900 python -m pytest tests/test_gpu_zy_sk_gemm.py -x -q -m gpu -k "ns_gemm" > $O/pytest_ns.log 2>&1; echo "exit $?" >> $O/pytest_ns.log)
grep -v "^  File\|^Extension\|^$" $O/pytest_ns.log | tail -15
(timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_lm.py -x -q -m gpu > $O/pytest_lm.log 2>&1; echo "exit $?" >> $O/pytest_lm.log)
grep -v "^  File\|^Extension\|^$" $O/pytest_lm.log | tail -15
(timeout 300 python tools/dep_trace.py --B 104 > $O/dep_trace_v2_b104.json 2> $O/dep_trace_v2_b104.err; echo "exit $?" >> $O/dep_trace_v2_b104.err)
(timeout 300 python tools/dep_trace.py --B 8 > $O/dep_trace_v2_b8.json 2> $O/dep_trace_v2_b8.err)
(B200_DEP_FUSED=2 timeout 300 python tools/dep_trace.py --B 1 > $O/dep_trace_v2_b1.json 2> $O/dep_trace_v2_b1.err)
head -c 1200 $O/dep_trace_v2_b104.json; tail -3 $O/dep_trace_v2_b104.err
(timeout 600 python tools/kbench.py --what gemm --M 48,104 --ns-sweep --only temporal > $O/kbench_ns.jsonl 2> $O/kbench_ns.err; echo "exit $?" >> $O/kbench_ns.err)
(timeout 300 python tools/kbench.py --what gemm --M 48,104 --ns-sweep --only text_linear >> $O/kbench_ns.jsonl 2>> $O/kbench_ns.err)
(timeout 300 python tools/kbench.py --what gemm --M 48,104 --ns-sweep --only depformer_in_all >> $O/kbench_ns.jsonl 2>> $O/kbench_ns.err)
tail -3 $O/kbench_ns.err
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-secondary > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['lm_step']['ms'],d['lm_b1'], d['roofline_gemm'])"
B200_DEP_FUSED=2 timeout 300 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline --skip-secondary --sessions 8 > $O/bench_b8_fusedb1.json 2> $O/bench_b8.err
python -c "
import json;d=json.loads(open('$O/bench_b8_fusedb1.json').read().strip().splitlines()[-1]);print('depfused=2:',d['ms_per_step'],d['lm_step']['ms'],d['lm_b1'])"
