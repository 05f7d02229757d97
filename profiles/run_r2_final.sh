# round 2, final GPU call on one GPU: the whole GPU suite, smoke, the default bench line and the reference arm
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/r2f
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -3 $O/smoke.log
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > $O/bench_reference.json 2> $O/bench_reference.err
timeout 300 python bench.py --workload cyber_c3 --parallel-envs 4194304 --skip-other-workloads > $O/cyber_c3_4194304.json 2>$O/cy.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:cyber_step_tiled -s 12 -c 1 -o $O/cyber_c3_4194304 python bench.py --workload cyber_c3 --parallel-envs 4194304 --skip-other-workloads --windows 1 > $O/ncu_cy.log 2>&1
ls -la $O
