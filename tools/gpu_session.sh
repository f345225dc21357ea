#!/usr/bin/env bash
# tools/gpu_session.sh — the standard evidence runs of a round, meant for `gpurun --timeout N -- 'bash tools/gpu_session.sh <stage>...'`.
# Every stage runs under its own `timeout`, writes under gpurun_out/ (scratch; copy what is to be judged into profiles/) and
# never prints a bench value from a run under ncu. Stages:
#   tests     full `-m gpu` suite
#   bench     default bench line (C3, 1 GPU) and the reference arm
#   launches  ncu launch list of one C3 solve (per-launch device times; compare SHARES with the bench line's kernel_classes)
#   spmv      ncu --set full of the dominant kernels (count-level forward / adjoint stream kernels)
#   gram      C'C + tssvd timing at C3 and ncu --set full of gram_adj_kernel
#   knn       kNN timing at 1,306,127 x 10 (exact-width instance against the multiple-of-8 one) and x 50; ncu of knn_kernel
#   sanitize  compute-sanitizer memcheck + racecheck over tools/sanitize_case.py (every kernel family, tiny sizes)
#   snn       Jaccard / SNN timing at 1,306,127 x 10, k = 20 (hash and general paths); ncu of snn_enumerate_kernel
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BENCH="python bench.py --config C3 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
for stage in "$@"; do
  case "$stage" in
    tests)
      timeout 600 python -m pytest tests -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log ;;
    bench)
      timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.json
      timeout 900 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 400 gpurun_out/bench_ref.json ;;
    launches)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/launches.log 2>&1
      wc -l gpurun_out/launches.csv ;;
    spmv)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fwd_stream_kernel|adj_stream_kernel" -s 20 -c 4 \
        -f -o gpurun_out/spmv $BENCH > gpurun_out/spmv.log 2>&1; ls -la gpurun_out/spmv.ncu-rep ;;
    gram)
      timeout 600 python tools/gram_check.py C3 > gpurun_out/gram_c3.log 2>&1; tail -1 gpurun_out/gram_c3.log
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_adj_kernel -s 1 -c 1 -f -o gpurun_out/gram \
        python tools/gram_check.py C2 > gpurun_out/gram_ncu.log 2>&1; ls -la gpurun_out/gram.ncu-rep ;;
    knn)
      KNN_VARIANTS=widths timeout 300 python tools/knn_check.py 1306127 10 20 > gpurun_out/knn_1306127_10.log 2>&1; tail -1 gpurun_out/knn_1306127_10.log
      KNN_VARIANTS=widths timeout 600 python tools/knn_check.py 1306127 50 20 > gpurun_out/knn_1306127_50.log 2>&1; tail -1 gpurun_out/knn_1306127_50.log
      KNN_VARIANTS=default timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_kernel -c 1 -f -o gpurun_out/knn \
        python tools/knn_check.py 131072 50 20 > gpurun_out/knn_ncu.log 2>&1; ls -la gpurun_out/knn.ncu-rep ;;
    snn)
      timeout 600 python tools/snn_check.py 1306127 10 20 > gpurun_out/snn_1306127.log 2>&1; tail -1 gpurun_out/snn_1306127.log
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:snn_enumerate_kernel -c 2 -f -o gpurun_out/snn \
        python tools/snn_check.py 262144 10 20 > gpurun_out/snn_ncu.log 2>&1; ls -la gpurun_out/snn.ncu-rep ;;
    sanitize)
      # memcheck (out-of-bounds / misaligned / leaks of device memory) and racecheck (shared-memory hazards) over a small pass
      # through every kernel family; the logs are the evidence kept under profiles/
      timeout 1500 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 3 python tools/sanitize_case.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
      timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_case.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck.log ;;
    *) echo "unknown stage: $stage" ;;
  esac
done
