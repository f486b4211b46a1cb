timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "peer_exchange or virtual" 2>&1 | tail -3
MFAR_NCU_RANGE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 3 --warmup 3 --others amazon_full:64:512 --cpu-budget-s 0 > gpurun_out/r2n_under_ncu.json 2> /dev/null
timeout 150 python tools/fuzz_parity.py --mode kernels --seconds 100 --seed 21 --out gpurun_out/r2n_fuzz_kernels.json 2>&1 | tail -3
timeout 120 python tools/fuzz_parity.py --mode api --seconds 70 --seed 22 --out gpurun_out/r2n_fuzz_api.json 2>&1 | tail -3
timeout 100 python tools/fuzz_parity.py --mode bm25 --seconds 50 --seed 23 --out gpurun_out/r2n_fuzz_bm25.json 2>&1 | tail -3
