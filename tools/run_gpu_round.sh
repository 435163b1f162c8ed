mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -m gpu --timeout 600 -k "full_size" 2>&1 | tail -30 > gpurun_out/t_all.log
tail -n 30 gpurun_out/t_all.log | cut -c1-250
