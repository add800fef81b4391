mkdir -p gpurun_out/r02b
cd /root/repo
O=gpurun_out/r02b
(timeout 600 python -m pytest tests/test_gpu_lm.py -x -q -m gpu > $O/pytest_lm.log 2>&1; echo "exit $?" >> $O/pytest_lm.log)
tail -5 $O/pytest_lm.log
(timeout 300 python tools/dep_trace.py --B 104 > $O/dep_trace_v2_b104.json 2> $O/dep_trace_v2_b104.err; echo "exit $?" >> $O/dep_trace_v2_b104.err)
(timeout 300 python tools/dep_trace.py --B 8 > $O/dep_trace_v2_b8.json 2> $O/dep_trace_v2_b8.err)
(B200_DEP_FUSED=2 timeout 300 python tools/dep_trace.py --B 1 > $O/dep_trace_v2_b1.json 2> $O/dep_trace_v2_b1.err)
(B200_DEP_KERNEL=1 timeout 300 python tools/dep_trace.py --B 104 > $O/dep_trace_v1_b104.json 2> $O/dep_trace_v1_b104.err)
(B200_DEP_KERNEL=1 timeout 300 python tools/dep_trace.py --B 8 > $O/dep_trace_v1_b8.json 2> $O/dep_trace_v1_b8.err)
cat $O/dep_trace_v2_b104.json | head -c 1500; tail -3 $O/dep_trace_v2_b104.err
(timeout 900 python -m pytest tests/test_gpu_zt_7b_parity.py tests/test_gpu_zw_frame_service.py tests/test_gpu_zu_other_configs.py -x -q -m gpu > $O/pytest_rest.log 2>&1; echo "exit $?" >> $O/pytest_rest.log)
tail -5 $O/pytest_rest.log
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-secondary > $O/bench_v2.json 2> $O/bench_v2.err
python -c "
import json;d=json.loads(open('$O/bench_v2.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['lm_step']['ms'],d['lm_b1'])"
