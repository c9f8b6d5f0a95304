#!/bin/bash
# multistep device path: the new parity test, the full GPU suite, then the 4 M clustered box at activeRung 0 / 2 / 4
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multistep" 2>&1 | tail -15 | tee gpurun_out/pytest_multistep.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r01i.log
for r in 0 2 4; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 --large-kind clustered --large-active-rung $r \
    > gpurun_out/bench_multistep_r$r.json 2> gpurun_out/bench_multistep_r$r.err
  python -c "import json;j=json.load(open('gpurun_out/bench_multistep_r$r.json'));print(json.dumps(j['large_box']))"
done
