#!/bin/bash
# round-2 final evidence: tests, bench, launch list of one 32-pair launch sequence, ncu --set full of every se3et kernel family
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --profile-from-start off -k regex:"kpconv_rows|kpconv_fused|kpconv_lift|gemm_tma|gemm_stream|gnstats_stream|gram_kernel|geo_embed_lookup|radius_cell|radius_query|flash_attention|groupnorm_double|maxpool|spm_|add_layernorm|anchor_max|upsample_concat|part_" -c 330 -o /tmp/prof_all python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kpconv_rows_kernel" -c 2 -o gpurun_out/prof_rows_final python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_rows_final.log 2>&1; echo "ncu3 rc=$?"
ls -la gpurun_out | head -30
