#!/bin/bash
# GPU-box script, round 2 second pass (1 GPU): seam tests again, ncu launch list + --set full of the re-tiled
# 16-bit stencil, kernel variants, fermion force A/B, deflation, compute-sanitizer.
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_seam.py tests/test_gpu_parity.py -x -q -m gpu -k "seam or congrad_matches_oracle or mixed_precision_congrad" > gpurun_out/pytest_${tag}.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
# launch list of the default bench command (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 600 --csv \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras \
    > gpurun_out/bench_under_ncu_${tag}.log 2>&1
# --set full of every stencil variant
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:dslash_half_kernel|dslash_kernel<float|dslash_kernel<double, 0" -c 8 -f \
    -o /tmp/prof_dslash_${tag} python profiles/prof_target2.py > gpurun_out/prof_target_${tag}.log 2>&1
ncu -i /tmp/prof_dslash_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_dslash_${tag}_raw.csv 2>/dev/null
ncu -i /tmp/prof_dslash_${tag}.ncu-rep --page details --csv > gpurun_out/prof_dslash_${tag}_details.csv 2>/dev/null
# kernel variants (register caps, no-math / no-link-load probes)
timeout 900 python profiles/variant_probe.py > gpurun_out/variant_probe_${tag}.log 2>&1
cp gpurun_out/variant_probe.json gpurun_out/variant_probe_${tag}.json 2>/dev/null
# partitioned stencil on one GPU as its own neighbour, at the 8-GPU local volume of configs[3]
timeout 600 python profiles/halo_probe.py > gpurun_out/halo_probe_${tag}.json 2> gpurun_out/halo_probe_${tag}.err
B200KS_FORCE_PARTITION=zt timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv \
    --log-file gpurun_out/launches_selfpartition_${tag}.csv python profiles/prof_selfpart.py > gpurun_out/prof_selfpart_${tag}.log 2>&1
# fermion force: fused backward staple kernel vs four small kernels
for split in 0 1; do
  B200KS_FORCE_SPLIT=$split timeout 600 python bench.py --workload force --steps 3 --no-cpu-baseline \
      > gpurun_out/bench_force_${tag}_split${split}.json 2> gpurun_out/bench_force_${tag}_split${split}.err
  B200KS_FORCE_SPLIT=$split timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_force_${tag}_split${split}.csv python profiles/prof_force.py > gpurun_out/prof_force_${tag}_split${split}.log 2>&1
done
# deflation: first numbers (500 vectors = 50 GB resident)
timeout 600 python bench.py --workload deflate --nvecs 500 --steps 5 > gpurun_out/bench_deflate_${tag}.json 2> gpurun_out/bench_deflate_${tag}.err
NVECS=64 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_deflate_${tag}.csv python profiles/prof_deflate.py > gpurun_out/prof_deflate_${tag}.log 2>&1
# compute-sanitizer on the halo / flag protocol (one GPU as its own neighbour) and on two members sharing the device
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu \
    -k "forced_self_partition and zt-dims3" > gpurun_out/sanitizer_memcheck_selfpartition_${tag}.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_memcheck_selfpartition_${tag}.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu \
    -k "forced_self_partition and zt-dims3" > gpurun_out/sanitizer_racecheck_selfpartition_${tag}.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_racecheck_selfpartition_${tag}.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_seam.py -q -m gpu \
    -k "multi_gpu_context_matches_oracle and 2-dims4" > gpurun_out/sanitizer_memcheck_multictx_${tag}.log 2>&1
echo "rc=$?" >> gpurun_out/sanitizer_memcheck_multictx_${tag}.log
tail -n 3 gpurun_out/pytest_${tag}.log
tail -n 2 gpurun_out/sanitizer_*_${tag}.log
