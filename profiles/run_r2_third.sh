# round 2: a host-path iteration -- parity first (short timeouts), then timing
cd "$(dirname "$0")/.."
O=gpurun_out/r2c
mkdir -p $O
( timeout 400 python -m pytest tests/test_host_step_gpu.py -x -q ) > $O/pytest_new.log 2>&1; tail -4 $O/pytest_new.log
timeout 300 python bench.py --skip-other-workloads > $O/c4_i8.json 2> $O/c4_i8.err || tail -3 $O/c4_i8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c/c4_i8.json').read().strip().splitlines()[-1])
for k in ('e2e','e2e_i16_actions','e2e_i8_actions','e2e_full_obs'):
    print(k, '%.4e'%d[k]['value'], '%.1f us'%(1e3*d[k]['ms_per_step']), d[k]['h2d_bytes_per_step'])
PY
