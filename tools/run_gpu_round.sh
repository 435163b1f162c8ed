mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -25 > gpurun_out/t_all.log
tail -n 25 gpurun_out/t_all.log | cut -c1-300
