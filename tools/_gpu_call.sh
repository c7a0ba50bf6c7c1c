cd $GRAFT_REPO_ROOT
timeout 300 python tools/profile_step.py egnn_20kp 8 100 bf16x3 > gpurun_out/r02_final_prof_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r02_final_prof_plain.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 3000 -c 400 --csv --log-file gpurun_out/r02_final_egnn20kp_launches.csv python tools/profile_step.py egnn_20kp 60 100 bf16x3 > gpurun_out/r02_final_ncu_list_egnn.log 2>&1; echo "rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 2000 -c 400 --csv --log-file gpurun_out/r02_final_gvp20kp_launches.csv python tools/profile_step.py gvp_20kp 110 100 bf16x3 > gpurun_out/r02_final_ncu_list_gvp.log 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:egnn_edge_ws -s 12 -c 1 -o gpurun_out/r02_final_egnn_edge_warm python tools/profile_step.py egnn_20kp 8 100 bf16x3 > gpurun_out/r02_final_ncu_egnn.log 2>&1; echo "rc=$?"
timeout 400 ncu --set full --clock-control none --cache-control none --import-source on -k regex:tc_linear -s 40 -c 3 -o gpurun_out/r02_final_tc_linear_warm python tools/profile_step.py egnn_20kp 8 100 bf16x3 > gpurun_out/r02_final_ncu_tcl.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
