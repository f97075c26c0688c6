#!/bin/bash
# one parametrised GPU session script: bash tools/gpu_run.sh <step> ...   (steps run in order; output under gpurun_out/)
mkdir -p gpurun_out
for step in "$@"; do
case $step in
  fpn)      echo "== fpn + linear kernel tests"; timeout 900 python -m pytest tests/test_gpu_fpn.py tests/test_gpu_linear.py -q --timeout 600 2>&1 | tail -15 ;;
  decoder)  echo "== decoder tests"; timeout 1200 python -m pytest tests/test_gpu_decoder.py tests/test_gpu_fpn.py -x -q --timeout 600 2>&1 | tail -15 ;;
  suite)    echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log ;;
  smoke)    echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
  bench)    echo "== bench"; timeout 900 python bench.py > gpurun_out/r02_bench_n1.jsonl 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/r02_bench_n1.jsonl ;;
  benchref) echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null > gpurun_out/r02_bench_reference.jsonl; cut -c1-200 gpurun_out/r02_bench_reference.jsonl ;;
  batch2)   echo "== bench batch 2 on one GPU (what a rank of the N=8 strong run executes)"; timeout 600 python bench.py --batch 2 --no-e2e --no-cpu --no-ref-cuda --steps 10 --warmup 3 --no-graph 2>/dev/null | tee gpurun_out/r02_bench_batch2_n1.jsonl | cut -c1-330
            timeout 600 python bench.py --batch 2 --no-e2e --no-cpu --no-ref-cuda --steps 10 --warmup 3 2>/dev/null | tee -a gpurun_out/r02_bench_batch2_n1.jsonl | cut -c1-330 ;;
  decbench) echo "== decoder bench"; timeout 900 python tools/decoder_bench.py 2>&1 | tee gpurun_out/r02_decoder_bench.txt | tail -30 ;;
  fpnbench) echo "== fpn bench"; timeout 900 python tools/fpn_bench.py 2>&1 | tee gpurun_out/r02_fpn_bench.txt | tail -40 ;;
  bwdcells) echo "== backward: cell-strided phase 2 A/B"; timeout 900 python -m pytest tests/test_gpu_bwd_sorted.py -x -q --timeout 600 2>&1 | tail -3
            timeout 600 python tools/bwd_ab2.py --set cells --reps 5 2>&1 | tee gpurun_out/r02_bwd_cells_ab.txt
            timeout 600 python tools/bwd_phases.py --variants 0,13,14,15,16 2>&1 | tee gpurun_out/r02_bwd_cells_phases.txt ;;
  sass)     echo "== SASS census of the shipped default kernels"; python tools/sass_census.py > gpurun_out/r02_sass_main_kernels.txt 2>&1; head -40 gpurun_out/r02_sass_main_kernels.txt ;;
  ncusorted) echo "== ncu --set full: default forward + anchor-sorted backward (cfg 2 x 16)"
            python tools/ncu_target.py 16 > gpurun_out/ncu_plain_target.log 2>&1 && \
            ncu --set full --clock-control none --import-source on -k regex:msda_ -s 2 -c 2 -o gpurun_out/prof_r02_final -f python tools/ncu_target.py 16 > gpurun_out/ncu_full.log 2>&1
            echo "rc=$?"; tail -2 gpurun_out/ncu_full.log ;;
  launches) echo "== ncu launch list of the bench command"
            python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ref-cuda > gpurun_out/ncu_plain_bench.log 2>&1 && \
            ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-ref-cuda > gpurun_out/ncu_bench.log 2>&1
            echo "rc=$?" ;;
  geoab)    echo "== forward: geometry warps with 16 / 20 / 24 / 28 consumer warps"; timeout 600 python tools/geo_ab.py 16 5 2>&1 | tee gpurun_out/r02_fwd_geo_warps_ab.txt ;;
  modbench) echo "== module / encoder benches"; timeout 600 python tools/module_bench.py 2>&1 | tee gpurun_out/r02_module_bench.txt | tail -12
            timeout 900 python tools/encoder_bench.py 2>&1 | tee gpurun_out/r02_encoder_bench.txt | tail -8 ;;
  allcfg)   echo "== bench lines of the other BASELINE configs"; rm -f gpurun_out/r02_bench_lines_final.jsonl
            for c in 1 3 4 5; do timeout 600 python bench.py --cfg $c --no-e2e --steps 10 --warmup 3 2>/dev/null | tee -a gpurun_out/r02_bench_lines_final.jsonl | cut -c1-200; done ;;
  *) echo "unknown step $step" ;;
esac
done
