mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 600 -x -k "window or joint or golden" 2>&1 | tail -3 > gpurun_out/t_all.log
timeout 300 python tools/attn_sweep.py > gpurun_out/attn_sweep.log 2>&1
tail -n 2 gpurun_out/t_all.log; cat gpurun_out/attn_sweep.log
