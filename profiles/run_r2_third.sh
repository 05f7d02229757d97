# round 2: a kernel / host-path iteration -- parity first (short timeouts), then timing
cd "$(dirname "$0")/.."
O=gpurun_out/r2c
mkdir -p $O
( timeout 400 python -m pytest tests/test_host_step_gpu.py -x -q ) > $O/pytest_new.log 2>&1; tail -4 $O/pytest_new.log
timeout 300 python bench.py --skip-other-workloads > $O/c4_rounds.json 2> $O/c4_rounds.err || tail -3 $O/c4_rounds.err
timeout 300 python bench.py --skip-other-workloads --parallel-envs 262144 > $O/c4_262144_rounds.json 2> $O/c4b.err || tail -3 $O/c4b.err
python profiles/time_host_step.py > $O/host_step.log 2>&1; tail -12 $O/host_step.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2c/c4*_rounds.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split('/')[-1], 'value %.3e'%d['value'], 'e2e %.3e (%.1f us)'%(d['e2e']['value'],1e3*d['e2e']['ms_per_step']), 'i16 %.3e (%.1f us)'%(d['e2e_i16_actions']['value'],1e3*d['e2e_i16_actions']['ms_per_step']), 'full %.3e'%d['e2e_full_obs']['value'])
PY
