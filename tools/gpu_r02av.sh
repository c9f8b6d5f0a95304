#!/bin/bash
# last GPU call of the round: the whole GPU suite (no -x: every test reports), incl. the CUDA path against the
# golden vectors of the reference's own CPU gravity
mkdir -p gpurun_out
timeout 45 python -m pytest tests -m gpu -q -rf 2>&1 | tail -12 | tee gpurun_out/r02av_pytest_gpu.log
