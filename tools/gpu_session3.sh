#!/bin/bash
# Third GPU session: new variants (256-bit forward, merged scatter), hybrid RED microbench, ncu of the RED microbench.
mkdir -p gpurun_out
echo "== pytest (variants + full-size)"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "variants or full_size or golden_f32" > gpurun_out/pytest_gpu3.log 2>&1; tail -8 gpurun_out/pytest_gpu3.log
echo "== sweep"; timeout 900 python tools/sweep.py --tag r01b > gpurun_out/sweep3.txt 2>&1; cat gpurun_out/sweep3.txt
echo "== microbench hybrid"; timeout 300 ./bm2f_b200/msda_microbench > gpurun_out/microbench3.txt 2>&1; grep -E "hybrid|tma|red_8xv4" gpurun_out/microbench3.txt
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -2 gpurun_out/bench3.err; python -c "
import json; d=json.load(open('gpurun_out/bench3.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['gather'], d['e2e'] and d['e2e']['value'], d['reference_cuda'] and d['reference_cuda']['speedup_ours'])"
echo "== ncu microbench red"
./bm2f_b200/msda_microbench > gpurun_out/mb_plain.log 2>&1 && \
ncu --metrics lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,lts__t_sectors_op_red.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"red_kernel|gather_kernel|hybrid" -c 60 --csv --log-file gpurun_out/ncu_microbench.csv ./bm2f_b200/msda_microbench > gpurun_out/ncu_mb.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_microbench.csv | cut -c1-300
