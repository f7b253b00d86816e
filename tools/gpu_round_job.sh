set -x
python -m pytest tests -m gpu -x -q > gpurun_out/test11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test11.log; tail -3 gpurun_out/test11.log
python bench.py > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_launches_1e7halos.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l4.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"los_nodes|paint_kernel" -c 8 -o gpurun_out/prof_r1d python bench.py --halos 1000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f4.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -8
