# end-of-session GPU job: parity suite, then the default bench (run with: gpurun -- 'bash tools/gpu_round_job.sh NN')
set -x
N=${1:-15}
python -m pytest tests -m gpu -x -q > gpurun_out/test$N.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test$N.log; tail -3 gpurun_out/test$N.log
python bench.py > gpurun_out/bench$N.json 2> gpurun_out/bench$N.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench${N}_ref.json 2> gpurun_out/bench${N}_ref.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
