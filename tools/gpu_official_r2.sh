#!/bin/bash
# round-2 artefacts on ONE GPU: GPU tests, smoke, bench (both arms), kernel microbench, ncu launch lists + full captures
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err
timeout 900 python bench.py --steps 100 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err
timeout 900 python tools/bench_kernels.py --out gpurun_out/kernels.json > gpurun_out/kernels.log 2>&1
RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 64 > gpurun_out/tgt_phase_64.log 2>&1
RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/tgt_phase_profile.py 512 > gpurun_out/tgt_phase_512.log 2>&1
# launch list + full capture of the dominant kernel on the bench command itself (64 panels, headline only)
# the launch form K4 tuned itself to in the plain run is fixed for the runs under ncu (timing candidates under a profiler means nothing)
FORM=$(python -c "
import json
f=json.loads([l for l in open('gpurun_out/bench.log') if l.startswith('{')][-1])['roofline'].get('launch_form') or {}
print('%d,%d' % (f.get('cluster', 0), f.get('sync_every', 0)))")
echo "K4 launch form (cluster,sync_every): $FORM"
PROF="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --skip-extra --k4-lockstep $FORM"
timeout 300 $PROF > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $PROF > gpurun_out/ncu_launches.log 2>&1
timeout 300 $PROF > gpurun_out/prof_plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:roi_pool_slice -s 3 -c 1 -o gpurun_out/prof_pool $PROF > gpurun_out/ncu_pool.log 2>&1
# the training-side kernels
P="python tools/prof_train.py 64"
timeout 300 $P > gpurun_out/prof_train_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv $P > gpurun_out/ncu_launches_train.log 2>&1
timeout 300 $P > gpurun_out/prof_train_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'rpn_targets_kernel|rpn_subsample|rpn_losses|roi_targets|select_samples|class_losses|match_candidates|class_ap' -s 16 -c 8 -o gpurun_out/prof_train $P > gpurun_out/ncu_train.log 2>&1
# decode / sort+NMS on 512 panels, detection kernels
P="python tools/prof_small.py 512"
timeout 300 $P > gpurun_out/prof_small_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'decode_clip|sort_nms' -s 2 -c 2 -o gpurun_out/prof_small $P > gpurun_out/ncu_small.log 2>&1
P="python tools/prof_detect.py"
timeout 300 $P > gpurun_out/prof_detect_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_detect.csv $P > gpurun_out/ncu_launches_detect.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_reference.log; cut -c1-1500 gpurun_out/bench.log; tail -3 gpurun_out/bench.err; tail -2 gpurun_out/ncu_pool.log; tail -2 gpurun_out/ncu_train.log
