set -x
cd /root/repo
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -3
REED_TMA_EPI=15 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gate_residual or epilogues" 2>&1 | tail -5
REED_TMA_EPI=15 REED_GATERES_SLOTS=3 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gate_residual" 2>&1 | tail -3
python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02z_gr_lsu.txt 2>&1
for R in 3 4 6 8; do REED_TMA_EPI=15 REED_GATERES_SLOTS=$R REED_GATERES_MIN_STAGES=2 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02z_gr_R$R.txt 2>&1; done
python profiles/bench_gemm.py --only "gate+res" --bn 256 > gpurun_out/r02z_gr_lsu_bn256.txt 2>&1
REED_TMA_EPI=15 python profiles/bench_gemm.py --only "gate+res" --bn 256 > gpurun_out/r02z_gr_R6_bn256.txt 2>&1
REED_TMA_EPI=15 REED_GEMM_DEBUG=3 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02z_gr_R6_epionly.txt 2>&1
REED_GEMM_DEBUG=3 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02z_gr_lsu_epionly.txt 2>&1
ncu --metrics gpu__time_duration.sum,launch__cluster_size --clock-control none --csv --log-file gpurun_out/r02z_cublas_kernels.csv python profiles/cublas_kernels.py > /dev/null 2>&1
tail -n +1 gpurun_out/r02z_gr_*.txt
