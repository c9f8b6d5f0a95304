#!/bin/bash
# tiny and degenerate particle sets through cb200_step_run
mkdir -p gpurun_out
timeout 60 python tools/edge_probe.py > gpurun_out/r02au_edge_probe.jsonl 2> gpurun_out/r02au_edge_probe.err
echo rc=$?
cat gpurun_out/r02au_edge_probe.jsonl
tail -5 gpurun_out/r02au_edge_probe.err
