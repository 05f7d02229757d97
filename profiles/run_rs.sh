# rideshare kernel iteration: parity tests, then the kernel's time at the saturating and the named batch size
cd "$(dirname "$0")/.."
O=gpurun_out/rs
mkdir -p $O; rm -f $O/*.json
( FRZ_RIDESHARE_KERNEL=tiles timeout 900 python -m pytest tests/test_rideshare_gpu.py tests/test_host_step_gpu.py -x -q ) > $O/pytest.log 2>&1; tail -5 $O/pytest.log
for rows in 8; do
  FRZ_RIDESHARE_KERNEL=tiles FRZ_RIDESHARE_TILE_ROWS=$rows timeout 300 python bench.py --workload rideshare_c2 --parallel-envs 524288 --skip-other-workloads --windows 3 > $O/tiles_r$rows.json 2> $O/tiles_r$rows.err || tail -3 $O/tiles_r$rows.err
done
for b in $RS_SWEEP; do
  for k in tiles groups; do
    FRZ_RIDESHARE_KERNEL=$k timeout 300 python bench.py --workload rideshare_c2 --parallel-envs $b --skip-other-workloads --windows 3 > $O/${k}_b$b.json 2> $O/${k}_b$b.err || tail -3 $O/${k}_b$b.err
  done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/rs/*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f.split('/')[-1], 'value %.3e'%d['value'], 'kernel_us %.1f'%(1e3*r['kernel_ms']), 'eager %.1f'%(1e3*r['kernel_ms_eager_launch']), 'frac %.3f'%r['frac'], 'rows %.2f'%r['mean_tasks_per_env'], 'e2e %.3e'%d['e2e']['value'])
    except Exception as e: print(f,'ERR',e)
PY
if [ -n "$RS_NCU" ]; then
  FRZ_RIDESHARE_KERNEL=tiles ncu --set full --clock-control none --import-source on -k regex:rideshare_tile -s 12 -c 1 -o $O/tile_$RS_NCU python bench.py --workload rideshare_c2 --parallel-envs 524288 --skip-other-workloads --windows 1 > $O/ncu_tile.log 2>&1; tail -2 $O/ncu_tile.log
fi
