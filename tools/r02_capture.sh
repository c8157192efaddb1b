#!/bin/bash
# Round-2 evidence capture (run under gpurun, one GPU): bench line, ncu launch list, ncu --set full of the hot kernels,
# the same for the opt-in TMA-staged RoIAlign path, compute-sanitizer over the new kernels.  Outputs: gpurun_out/r02_*.
set -u
O=gpurun_out
timeout -s KILL 400 python bench.py --steps 30 --warmup 5 > $O/r02_bench.json 2> $O/r02_bench.err
timeout -s KILL 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/r02_bench_reference_arm.json 2>> $O/r02_bench.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 > $O/r02_bench_under_ncu.log 2>&1
timeout -s KILL 500 ncu --set full --clock-control none --import-source on \
  -k regex:"roi_gather|roi_prologue|iou_tile|iou_exact|nms_mask|nms_exact|nms_scan|split_count|split_scatter|pack_detections|feature_refine|align_conv_tc|rec_kernel" \
  -o $O/r02_prof python tools/prof_ops.py --reps 1 > $O/r02_ncu.log 2>&1
JDET_ROI_TMA=1 timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"roi_gather|roi_prologue" \
  -o $O/r02_prof_tma python tools/prof_ops.py --reps 1 --only roi > $O/r02_ncu_tma.log 2>&1
JDET_ROI_TMA=1 timeout -s KILL 100 python tools/roi_time.py 2 > $O/r02_roi_time_tma.txt 2>&1
timeout -s KILL 100 python tools/roi_time.py 2 > $O/r02_roi_time_default.txt 2>&1
timeout -s KILL 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu \
  -k "tma_staged_path or nms_record or nms_poly or candidate_queue or horizontal or feature_refine or ml_nms_bit_exact or strict_threshold" > $O/r02_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> $O/r02_sanitizer_memcheck.log
tail -3 $O/r02_sanitizer_memcheck.log
tail -2 $O/r02_ncu.log
