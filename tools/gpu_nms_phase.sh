mkdir -p gpurun_out
bash tools/build_prof.sh > gpurun_out/build_prof.log 2>&1; tail -2 gpurun_out/build_prof.log
RADNET_NMS_CLUSTER=1 RADNET_B200_LIB=rock_art_radnet_b200/_C/libradnet_b200_prof.so timeout 300 python tools/nms_phase_profile.py 2>&1 | tail -6
