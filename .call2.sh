cd /root/repo
REED_TMA_EPI=15 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gate_residual or epilogues" > gpurun_out/r02za_tests.txt 2>&1
REED_TMA_EPI=15 REED_GATERES_SLOTS=2 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gate_residual" >> gpurun_out/r02za_tests.txt 2>&1
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" >> gpurun_out/r02za_tests.txt 2>&1
for R in 2 3; do REED_TMA_EPI=15 REED_GATERES_SLOTS=$R python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02za_gr_R$R.txt 2>&1; done
REED_TMA_EPI=15 python profiles/bench_gemm.py --only "gate+res" --bn 256 > gpurun_out/r02za_gr_R3_bn256.txt 2>&1
REED_TMA_EPI=15 REED_GEMM_DEBUG=3 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02za_gr_R3_epionly.txt 2>&1
REED_TMA_EPI=15 REED_GEMM_DEBUG=4 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02za_gr_noepi.txt 2>&1
grep -h "passed\|failed\|error" gpurun_out/r02za_tests.txt
tail -n +1 gpurun_out/r02za_gr_*.txt
