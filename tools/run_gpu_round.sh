mkdir -p gpurun_out
timeout 600 python tools/configs_bench.py --only backbones > gpurun_out/backbones_bench.jsonl 2> gpurun_out/backbones_bench.err
cut -c1-260 gpurun_out/backbones_bench.jsonl; tail -n 3 gpurun_out/backbones_bench.err
