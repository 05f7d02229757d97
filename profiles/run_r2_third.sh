# round 2: a kernel iteration -- parity first (short timeouts), then timing
cd "$(dirname "$0")/.."
O=gpurun_out/r2c
mkdir -p $O
( timeout 300 python -m pytest tests/test_cyber_gpu.py tests/test_host_step_gpu.py tests/test_philox_parity_gpu.py -x -q -k "cyber" ) > $O/pytest_new.log 2>&1; tail -3 $O/pytest_new.log
for w in "cyber_c3 16384" "cyber_c3 4194304"; do
  set -- $w
  timeout 240 python bench.py --workload $1 --parallel-envs $2 --skip-other-workloads --windows 3 > $O/$1_$2.json 2> $O/$1_$2.err || tail -3 $O/$1_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/cyber_c3_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.3e'%d['value'], 'kernel_us %.1f'%(1e3*r['kernel_ms']), 'eager %.1f'%(1e3*r['kernel_ms_eager_launch']), 'frac %.3f'%r['frac'], 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
