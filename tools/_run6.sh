mkdir -p gpurun_out/r02f
cd /root/repo
O=gpurun_out/r02f
(timeout 600 python -m pytest tests/test_gpu_mimi.py tests/test_gpu_zs_mimi_tc.py tests/test_gpu_zw_frame_service.py -x -q -m gpu > $O/pytest_mimi.log 2>&1; echo "exit $?" >> $O/pytest_mimi.log)
grep -v "^  File\|^Extension\|^$" $O/pytest_mimi.log | tail -6
timeout 300 python tools/kbench.py --what mimi --B 1,16,104 > $O/kbench_mimi_pdl.jsonl 2> $O/kbench.err
B200_MIMI_PDL=0 timeout 300 python tools/kbench.py --what mimi --B 1,16,104 > $O/kbench_mimi_nopdl.jsonl 2>> $O/kbench.err
cat $O/kbench_mimi_pdl.jsonl $O/kbench_mimi_nopdl.jsonl
timeout 400 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-secondary > $O/bench.json 2> $O/bench.err
python -c "
import json;d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]);print(d['ms_per_step'],d['e2e']['ms_per_step'],d['lm_step']['ms'],d['mimi'])"
