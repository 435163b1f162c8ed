mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_v16.json 2> gpurun_out/bench_v16.err
python -c "
import json; d=json.load(open('gpurun_out/bench_v16.json')); print(json.dumps(d.get('hbm_kernels'), indent=1)); print(d['value'], d['roofline']['frac'], d['cpu_baseline'])"
tail -n 3 gpurun_out/bench_v16.err
