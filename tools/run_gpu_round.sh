mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -8 > gpurun_out/t_all.log
timeout 300 python tools/step_timeline.py --conv resnet101 > gpurun_out/timeline_resnet101.log 2>&1
tail -n 8 gpurun_out/t_all.log; head -n 14 gpurun_out/timeline_resnet101.log
