# round 2, first GPU call: tests, the default bench line, the reference arm, launch list + one full capture per kernel
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/r2
mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
( time python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
python bench.py --skip-other-workloads --steps 20 --warmup 5 > $O/plain_c4.json 2>$O/plain_c4.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_wildfire_c4.csv python bench.py --skip-other-workloads --steps 20 --warmup 5 --windows 1 > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wildfire_step -s 12 -c 1 -o $O/wildfire_c4 python bench.py --skip-other-workloads --steps 20 --warmup 5 --windows 1 > $O/ncu_c4.log 2>&1
python bench.py --workload rideshare_c2 --parallel-envs 524288 --skip-other-workloads > $O/plain_rs524k.json 2>$O/plain_rs.err && \
ncu --set full --clock-control none --import-source on -k regex:rideshare_step -s 12 -c 1 -o $O/rideshare_c2_524288 python bench.py --workload rideshare_c2 --parallel-envs 524288 --skip-other-workloads --windows 1 > $O/ncu_rs.log 2>&1
ls -la $O
