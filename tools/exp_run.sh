OLD=$PWD/multifield-adaptive-retrieval_b200/mfar_b200/libmfar_b200_old.so
timeout 600 python -m pytest tests/test_gpu_parity_at_scale.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
for lib in old new; do
  if [ $lib = old ]; then export MFAR_LIB=$OLD; else unset MFAR_LIB; fi
  timeout 100 python tools/quick_bench.py --workload mag_full --batches 1,64,512 --iters 20 --tag $lib 2>&1 | grep -E "^\{|Error" | cut -c1-200
  timeout 100 python tools/quick_bench.py --workload scale_10m_all --docs 1250000 --batches 1,512 --iters 20 --tag $lib 2>&1 | grep -E "^\{|Error" | cut -c1-200
  timeout 100 python tools/quick_bench.py --workload scale_10m_single --batches 64,128,512 --iters 20 --tag $lib 2>&1 | grep -E "^\{|Error" | cut -c1-200
  timeout 100 python tools/quick_bench.py --workload amazon_full --batches 64,512 --iters 20 --tag $lib 2>&1 | grep -E "^\{|Error" | cut -c1-200
done
