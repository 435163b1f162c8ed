mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -12 > gpurun_out/t_all.log
tail -n 12 gpurun_out/t_all.log
