#!/bin/bash
# round 2, first box visit: parity (new p-p kernel, Ewald vs the reference kernel, ABI link test),
# p-p kernel A/B at 4 M and 256^3, ncu of the new kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02a_pytest_gpu.log
for v in 0 3 2; do
  CB200_PP_VARIANT=$v timeout 300 python bench.py --steps 50 --warmup 5 --large-n 4194304 --no-cpu-baseline \
    > gpurun_out/r02a_pp_v$v.json 2> gpurun_out/r02a_pp_v$v.err
  python - <<PY
import json
j = json.load(open("gpurun_out/r02a_pp_v$v.json"))
k, l = j["kernels"], j.get("large_box", {})
print("variant $v cube300: pp_ms %.4f pp/s %.3e pc_ms %.4f step %.4f | 4M: pp_ms %.3f pc_ms %.3f step %.2f phases %s" % (
    k["pp_ms"], k["pp_interactions_per_s"], k["pc_ms"], j["ms_per_step"], l.get("rank0_pp_ms", 0), l.get("rank0_pc_ms", 0),
    l.get("ms_per_step", 0), l.get("rank0_phases_ms")))
PY
done
timeout 600 python bench.py --steps 20 --warmup 5 --large-n 16777216 --no-cpu-baseline > gpurun_out/r02a_256.json 2> gpurun_out/r02a_256.err
python -c "
import json; l=json.load(open('gpurun_out/r02a_256.json'))['large_box']; print('256^3:', {k:l[k] for k in ('ms_per_step','pc_pairs','pp_pairs','rank0_phases_ms','rank0_pc_ms','rank0_pp_ms','rank0_ewald_ms','rank0_pc_frac_of_fp32_peak','rank0_hbm_in_use_gb')})"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"part_list" -s 6 -c 1 -f -o gpurun_out/r02a_prof_pp \
  python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --large-n 0 > gpurun_out/r02a_prof_pp.log 2>&1
ls -la gpurun_out | tail -4
