# round 2: a kernel iteration -- parity first (short timeouts), then timing
cd "$(dirname "$0")/.."
O=gpurun_out/r2c
mkdir -p $O
( timeout 300 python -m pytest tests/test_cyber_gpu.py tests/test_host_step_gpu.py tests/test_philox_parity_gpu.py -x -q -k "cyber" ) > $O/pytest_new.log 2>&1; tail -4 $O/pytest_new.log
for n in 2 3; do
  FRZ_CYBER_BUFFERS=$n timeout 240 python bench.py --workload cyber_c3 --parallel-envs 4194304 --skip-other-workloads --windows 3 > $O/cyber_buffers$n.json 2> $O/cyber_buffers$n.err || tail -3 $O/cyber_buffers$n.err
done
FRZ_CYBER_BUFFERS=3 timeout 240 python bench.py --workload cyber_c3 --parallel-envs 1048576 --skip-other-workloads --windows 3 > $O/cyber_1m_buffers3.json 2> $O/e.err
FRZ_CYBER_BUFFERS=2 timeout 240 python bench.py --workload cyber_c3 --parallel-envs 1048576 --skip-other-workloads --windows 3 > $O/cyber_1m_buffers2.json 2> $O/e.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/cyber_*buffers*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.3e'%d['value'], 'kernel_us %.1f'%(1e3*r['kernel_ms']), 'eager %.1f'%(1e3*r['kernel_ms_eager_launch']), 'frac %.3f'%r['frac'], 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
