mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -4 > gpurun_out/t_all.log
timeout 900 python bench.py > gpurun_out/bench_v15.json 2> gpurun_out/bench_v15.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_v15.json 2> gpurun_out/bench_ref_v15.err
N=$(python tools/profile_step.py --count-only | awk '/launches_per_step/{print $2}')
ncu --kernel-name-base demangled -k regex:mvlt:: --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/r01_launches_step_b64_v15.csv python tools/profile_step.py --passes 2 > gpurun_out/ncu_v15.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -n 2 gpurun_out/t_all.log; cut -c1-400 gpurun_out/bench_v15.json; cut -c1-300 gpurun_out/bench_ref_v15.json; tail -n 2 gpurun_out/smoke.log
