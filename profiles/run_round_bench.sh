set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/r1
for w in wildfire_c4 wildfire_c1 rideshare_c2 cyber_c3; do
  python bench.py --workload $w > gpurun_out/r1/bench_$w.json 2> gpurun_out/r1/bench_$w.err || tail -5 gpurun_out/r1/bench_$w.err
done
python bench.py --workload wildfire_c4 --parallel-envs 262144 --steps 30 > gpurun_out/r1/bench_wildfire_c4_262144.json 2>gpurun_out/r1/e1.err
python bench.py --workload wildfire_c1 --parallel-envs 524288 --steps 30 > gpurun_out/r1/bench_wildfire_c1_524288.json 2>gpurun_out/r1/e2.err
python bench.py --workload rideshare_c2 --parallel-envs 524288 --steps 30 > gpurun_out/r1/bench_rideshare_c2_524288.json 2>gpurun_out/r1/e3.err
python bench.py --workload cyber_c3 --parallel-envs 4194304 --steps 30 > gpurun_out/r1/bench_cyber_c3_4194304.json 2>gpurun_out/r1/e4.err
python profiles/time_host_step.py > gpurun_out/r1/host_step.log 2>&1
python bench.py --impl reference --steps 5 > gpurun_out/r1/bench_reference_wildfire_c4.json 2>gpurun_out/r1/e5.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r1/plain_c4.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1/launches_wildfire_c4.csv python bench.py --steps 20 --warmup 3 > gpurun_out/r1/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wildfire_step -s 12 -c 1 -o gpurun_out/r1/wildfire_c4 python bench.py --steps 20 --warmup 3 > gpurun_out/r1/ncu_c4.log 2>&1
python bench.py --workload rideshare_c2 --steps 20 --warmup 3 > gpurun_out/r1/plain_rs.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rideshare_step -s 12 -c 1 -o gpurun_out/r1/rideshare_c2 python bench.py --workload rideshare_c2 --steps 20 --warmup 3 > gpurun_out/r1/ncu_rs.log 2>&1
python bench.py --workload cyber_c3 --steps 20 --warmup 3 > gpurun_out/r1/plain_cy.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cyber_step_tiled -s 12 -c 1 -o gpurun_out/r1/cyber_c3 python bench.py --workload cyber_c3 --steps 20 --warmup 3 > gpurun_out/r1/ncu_cy.log 2>&1
python bench.py --workload wildfire_c1 --steps 20 --warmup 3 > gpurun_out/r1/plain_c1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wildfire_step -s 12 -c 1 -o gpurun_out/r1/wildfire_c1 python bench.py --workload wildfire_c1 --steps 20 --warmup 3 > gpurun_out/r1/ncu_c1.log 2>&1
for f in gpurun_out/r1/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get('roofline',{}); e=d.get('e2e',{}); c=d.get('cpu_baseline',{})
    print(sys.argv[1].split('/')[-1], 'value %.3e'%d['value'], 'kernel_us %.1f'%(1e3*r.get('kernel_ms',0)), 'eager_us %.1f'%(1e3*r.get('kernel_ms_eager_launch',0)), 'frac %.3f'%r.get('frac',0), 'e2e %.3e'%e.get('value',0), 'cpu %.3e'%c.get('value',0), d.get('clocks'))
except Exception as ex: print(sys.argv[1], 'ERR', ex)
PY
done
