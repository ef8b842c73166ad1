MP3B_RS_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_resample_tc' -s 2 -c 1 -f -o gpurun_out/rs_tc4 python tools/time_output_ops.py > gpurun_out/rs_tc_ncu.log 2>&1
echo rc=$?
