# round 2, second GPU call: the default bench line (with the other named workloads), rideshare at the saturating size
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/r2b
mkdir -p $O
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err
timeout 300 python bench.py --workload rideshare_c2 --parallel-envs 524288 --skip-other-workloads > $O/rideshare_c2_524288.json 2>$O/rs.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rideshare_tile -s 12 -c 1 -o $O/rideshare_c2_524288 python bench.py --workload rideshare_c2 --parallel-envs 524288 --skip-other-workloads --windows 1 > $O/ncu_rs.log 2>&1
ls -la $O
