mkdir -p gpurun_out
timeout 240 python tools/conv_diag.py > gpurun_out/conv_diag.log 2>&1; echo "diag rc=$?" >> gpurun_out/conv_diag.log
timeout 900 python -m pytest tests/test_resnet_gpu.py -q --timeout 300 -s 2>&1 | tail -60 > gpurun_out/t_resnet.log
timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 --ignore tests/test_resnet_gpu.py 2>&1 | tail -15 > gpurun_out/t_all.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_swin.json 2> gpurun_out/bench_swin.err
timeout 400 python bench.py --conv resnet101 --no-cpu-baseline > gpurun_out/bench_resnet101.json 2> gpurun_out/bench_resnet101.err
tail -5 gpurun_out/conv_diag.log gpurun_out/t_resnet.log gpurun_out/t_all.log; cat gpurun_out/bench_resnet101.json | cut -c1-400
