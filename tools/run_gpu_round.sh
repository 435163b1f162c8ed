mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -4 > gpurun_out/t_all.log
timeout 600 python tools/configs_bench.py --only config2 > gpurun_out/configs_bench2.jsonl 2> gpurun_out/configs_bench.err
tail -n 3 gpurun_out/t_all.log; cut -c1-200 gpurun_out/configs_bench2.jsonl
