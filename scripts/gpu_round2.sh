#!/bin/bash
# Round-2 evidence in ONE gpurun call (1 GPU): parity tests, smoke, both bench arms, ncu launch list of one training step,
# ncu --set full of the fused GSL kernel (B=32 batch launch and streaming size) and of the plane GEMM, compute-sanitizer.
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2_gpu.txt
timeout 600 python -m pytest tests -m gpu -q --maxfail=20 --tb=short -p no:cacheprovider > $OUT/r2_tests.log 2>&1; tail -2 $OUT/r2_tests.log | cut -c1-200
timeout 200 python __graft_entry__.py --smoke > $OUT/r2_smoke.log 2>&1; tail -1 $OUT/r2_smoke.log | cut -c1-200
timeout 900 python bench.py > $OUT/r2_bench.json 2> $OUT/r2_bench.err; echo "bench exit $? bytes $(wc -c < $OUT/r2_bench.json)"
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/r2_bench_ref.json 2> $OUT/r2_bench_ref.err; cut -c1-300 $OUT/r2_bench_ref.json
# launch list of one training step (issue order, ncu durations)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_step_launches.csv \
    python scripts/one_step.py snopes fp32 > $OUT/r2_step_launches.log 2>&1; echo "launch list rows $(wc -l < $OUT/r2_step_launches.csv)"
# full captures
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gather_row_kernel<\(bool\)1" -c 1 -f -o $OUT/r2_prof_gsl_b32 \
    python scripts/one_step.py snopes fp32 > $OUT/r2_ncu_gsl_b32.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:gather_row_kernel<\(bool\)1" -s 2 -c 1 -f -o $OUT/r2_prof_gsl_stream \
    python scripts/prof_gsl_stream.py > $OUT/r2_ncu_gsl_stream.log 2>&1
timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bp_kernel -c 8 -f -o $OUT/r2_prof_gemm \
    python scripts/one_step.py snopes fp32 > $OUT/r2_ncu_gemm.log 2>&1
ls -la $OUT/*.ncu-rep | cut -c1-150
# sanitizer: memcheck over the list kernels and a small model step; racecheck over the kernels that recycle shared memory
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_lists.py -q -x -p no:cacheprovider -k "fused_gsl_on_lists or texts_with_more or lists_match" > $OUT/r2_memcheck_lists.log 2>&1; tail -4 $OUT/r2_memcheck_lists.log | cut -c1-200
timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python __graft_entry__.py --smoke > $OUT/r2_memcheck_smoke.log 2>&1; tail -4 $OUT/r2_memcheck_smoke.log | cut -c1-200
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_lists.py -q -x -p no:cacheprovider -k "fused_gsl_on_lists and 4-30-300" > $OUT/r2_racecheck_lists.log 2>&1; tail -4 $OUT/r2_racecheck_lists.log | cut -c1-200
