mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 600 -k "gemm" 2>&1 | tail -4 > gpurun_out/t_all.log
timeout 300 python tools/gemm_graph_bench.py --flush --only fc2 > gpurun_out/gemm_shapes_sk.log 2>&1
timeout 300 python tools/gemm_graph_bench.py --flush --only fo >> gpurun_out/gemm_shapes_sk.log 2>&1
tail -n 3 gpurun_out/t_all.log | cut -c1-200; cat gpurun_out/gemm_shapes_sk.log
