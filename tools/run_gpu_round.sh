mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -q --timeout 300 2>&1 | tail -5 > gpurun_out/t_resnet.log
timeout 300 python tools/step_timeline.py --conv resnet101 > gpurun_out/timeline_resnet101.log 2>&1
tail -n 5 gpurun_out/t_resnet.log; head -n 40 gpurun_out/timeline_resnet101.log
