#!/usr/bin/env bash
# One-shot GPU validation bundle (run on a B200 box, e.g. `gpurun --timeout 600 -- 'bash tools/gpu_checks.sh'`).
# Everything lands under gpurun_out/; each step is bounded by its own timeout.  ~5 GPU-minutes in total:
#   1. pytest -m gpu                      (parity suite, ~25 s)
#   2. smoke()                            (driver's smoke entry)
#   3. randomised sweeps                  (kernels / api / bm25 / train, tools/fuzz_parity.py)
#   4. bench.py default line              (10M docs x 8 fields, Q=512 + Q=1/64)
#   5. ncu launch list of a short bench   (per-launch durations; compare SHARES with the bench line)
set -u
mkdir -p gpurun_out
SEED=${SEED:-11}
(timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 2>&1 | tail -20) > gpurun_out/checks_pytest.log 2>&1
tail -1 gpurun_out/checks_pytest.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6) > gpurun_out/checks_smoke.log 2>&1
tail -4 gpurun_out/checks_smoke.log
for mode in kernels api bm25 train; do
  (timeout 120 python tools/fuzz_parity.py --mode $mode --seconds ${FUZZ_SECONDS:-30} --seed $SEED \
      --out gpurun_out/checks_fuzz_$mode.json 2>&1 | tail -8) > gpurun_out/checks_fuzz_$mode.log 2>&1
  tail -1 gpurun_out/checks_fuzz_$mode.log
done
(timeout 300 python bench.py > gpurun_out/checks_bench.json 2> gpurun_out/checks_bench.err)
head -c 600 gpurun_out/checks_bench.json; echo
if [ "${NCU:-1}" = "1" ]; then
  (MFAR_NCU_RANGE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
      --csv --log-file gpurun_out/checks_launches.csv python bench.py --steps 3 --warmup 3 --extra-batches "" \
      --cpu-budget-s 0 > gpurun_out/checks_bench_under_ncu.json 2> gpurun_out/checks_ncu.err)
  tail -3 gpurun_out/checks_launches.csv
fi
