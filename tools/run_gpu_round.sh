mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -12 > gpurun_out/t_all.log
timeout 300 python tools/gemm_graph_bench.py --flush > gpurun_out/gemm_shapes_sk.log 2>&1
MVLT_STREAMK=0 timeout 300 python tools/gemm_graph_bench.py --flush > gpurun_out/gemm_shapes_nosk.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-roofline > gpurun_out/bench_swin.json 2> gpurun_out/bench_swin.err
tail -n 6 gpurun_out/t_all.log | cut -c1-200; paste <(awk '{print $1,$2,$3,$4,$5,$6}' gpurun_out/gemm_shapes_sk.log) <(awk '{print $3,$4,$5,$6}' gpurun_out/gemm_shapes_nosk.log); cut -c1-200 gpurun_out/bench_swin.json
