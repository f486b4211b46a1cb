ncu --set full --clock-control none --import-source on -k regex:score_qs -s 4 -c 2 -o gpurun_out/r2m_qs_10m_q512 -f python bench.py --steps 2 --warmup 3 --others none --extra-batches "" --cpu-budget-s 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_qs -s 4 -c 2 -o gpurun_out/r2m_qs_amazon_q512 -f python tools/quick_bench.py --workload amazon_full --batches 512 --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_tc -s 4 -c 2 -o gpurun_out/r2m_tc_amazon_q64 -f python tools/quick_bench.py --workload amazon_full --batches 64 --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_qs -s 4 -c 2 -o gpurun_out/r2m_qs_mag_q512 -f python tools/quick_bench.py --workload mag_full --batches 512 --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_tc -s 0 -c 1 -o gpurun_out/r2m_tc_10m_q1 -f python tools/quick_bench.py --workload scale_10m_all --batches 1 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
