cd /root/repo
python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gate_residual or epilogues" > gpurun_out/r02zb_tests.txt 2>&1
REED_GATERES_MAX_K=100000 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gate_residual or epilogues" >> gpurun_out/r02zb_tests.txt 2>&1
python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02zb_gr_default.txt 2>&1
REED_GATERES_MAX_K=100000 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02zb_gr_all.txt 2>&1
REED_GATERES_MAX_K=100000 REED_GATERES_SLOTS=3 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02zb_gr_all_R3.txt 2>&1
REED_GEMM_DEBUG=3 python profiles/bench_gemm.py --only "gate+res" > gpurun_out/r02zb_gr_epionly.txt 2>&1
python bench.py --steps 20 --skip-cpu > gpurun_out/r02zb_bench.json 2> gpurun_out/r02zb_bench.err
REED_TMA_EPI=7 python bench.py --steps 20 --skip-cpu > gpurun_out/r02zb_bench_lsu.json 2>> gpurun_out/r02zb_bench.err
grep -h "passed\|failed\|error" gpurun_out/r02zb_tests.txt
tail -n +1 gpurun_out/r02zb_gr_*.txt
python -c "
import json
for f in ('gpurun_out/r02zb_bench.json','gpurun_out/r02zb_bench_lsu.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
"
