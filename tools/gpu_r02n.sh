#!/bin/bash
mkdir -p gpurun_out
CB200_LIB=changa_b200/variants/stats.so timeout 300 python tools/step_probe.py --n 16777216 --steps 1 2>&1 >/dev/null | sort | uniq -c | tee gpurun_out/r02n_stats_256.log
CB200_LIB=changa_b200/variants/stats.so timeout 300 python tools/step_probe.py --n 4194304 --kind clustered --steps 1 2>&1 >/dev/null | sort | uniq -c | tee gpurun_out/r02n_stats_4Mclu.log
