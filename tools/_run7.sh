ncu --set full --clock-control none --import-source on -k regex:jq_traj_kernel -c 1 -f -o gpurun_out/r02_lat_cnot2_pipe python tools/ncu_target.py cnot2 1 > gpurun_out/ncu7a.log 2>&1
JQ_LAT_PIPE=0 ncu --set full --clock-control none --import-source on -k regex:jq_traj_kernel -c 1 -f -o gpurun_out/r02_lat_cnot2_nopipe python tools/ncu_target.py cnot2 1 > gpurun_out/ncu7b.log 2>&1
tail -2 gpurun_out/ncu7a.log gpurun_out/ncu7b.log
