mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_resnet_gpu.py -q --timeout 300 -k "stem or golden" 2>&1 | tail -5 > gpurun_out/t_resnet.log
ncu --set full --clock-control none --import-source on -k regex:resnet_stem -s 1 -c 1 -o gpurun_out/stem -f python tools/profile_step.py --conv resnet101 --passes 2 > gpurun_out/ncu_resnet.log 2>&1
tail -n 3 gpurun_out/t_resnet.log
