#!/bin/bash
# Round-end evidence on one B200 (run under gpurun from the repo root): bench records of the other workloads, the ncu
# launch list of a short bench, ncu --set full captures of the stream and tail kernels (each only behind a plain run of
# the same command, as /opt/skills/guides/B200_PROFILING.md asks).  tools/ncu_summary.py turns them into profiles/.
set -x
B="python bench.py"
$B --workload mel --no-config4 > gpurun_out/r2_mel_bench_8192.json 2> gpurun_out/r2_mel.err
$B --workload lfcc_ragged --no-config4 > gpurun_out/r2_ragged_bench_4096.json 2> gpurun_out/r2_ragged.err
$B --variant fft --no-config4 --no-e2e > gpurun_out/r2_fft_variant_bench_4096.json 2> gpurun_out/r2_fftv.err
$B --impl reference --steps 3 --warmup 1 > gpurun_out/r2_reference_arm.json 2> gpurun_out/r2_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-config4 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-config4 > gpurun_out/ncu_launches.log 2>&1
python tests/cuda/gemm_prof.py dft_gemm 1184 > gpurun_out/plain_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fe_stream -s 3 -c 1 -o gpurun_out/r2_final_stream python tests/cuda/gemm_prof.py dft_gemm 1184 > gpurun_out/ncu_stream.log 2>&1
python tests/cuda/tail_prof.py > gpurun_out/plain_tail.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fe_tail -s 2 -c 1 -o gpurun_out/r2_final_tail python tests/cuda/tail_prof.py > gpurun_out/ncu_tail.log 2>&1
tail -n 2 gpurun_out/plain_gemm.log; tail -n 2 gpurun_out/plain_tail.log
