# round 2: the small-grid wildfire kernel -- parity first (short timeouts), then timing
cd "$(dirname "$0")/.."
O=gpurun_out/r2c
mkdir -p $O
( timeout 420 python -m pytest tests/test_wildfire_gpu.py -x -q -k "not full_size" ) > $O/pytest_new.log 2>&1; tail -4 $O/pytest_new.log
( timeout 300 python -m pytest tests/test_philox_parity_gpu.py -x -q -k "wildfire and not full_size" ) > $O/pytest_philox.log 2>&1; tail -3 $O/pytest_philox.log
for w in "wildfire_c1 524288" "wildfire_5x6 262144"; do
  set -- $w
  timeout 240 python bench.py --workload $1 --parallel-envs $2 --skip-other-workloads --windows 3 > $O/$1_$2.json 2> $O/$1_$2.err || tail -3 $O/$1_$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/wildfire*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.3e'%d['value'], 'kernel_us %.1f'%(1e3*r['kernel_ms']), 'eager %.1f'%(1e3*r['kernel_ms_eager_launch']), 'frac %.3f'%r['frac'], 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
