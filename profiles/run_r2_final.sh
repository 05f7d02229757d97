# round 2, final GPU call on one GPU: the whole GPU suite, smoke, the default bench line
set -x
cd "$(dirname "$0")/.."
O=gpurun_out/r2f
mkdir -p $O
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/smoke.log 2>&1; tail -3 $O/smoke.log
( time timeout 900 python bench.py ) > $O/bench_default.json 2> $O/bench_default.err; tail -c 200 $O/bench_default.err
