mkdir -p gpurun_out
timeout 600 python tools/configs_bench.py > gpurun_out/configs_bench.jsonl 2> gpurun_out/configs_bench.err
cat gpurun_out/configs_bench.jsonl | cut -c1-330; tail -n 3 gpurun_out/configs_bench.err
