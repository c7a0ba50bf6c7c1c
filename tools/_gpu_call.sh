cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02e_pytest.log
timeout 200 python tools/eg_phase_times.py > gpurun_out/r02e_eg_phase.txt 2>&1; echo "rc=$?"; cat gpurun_out/r02e_eg_phase.txt
KPD_GVP_SERIAL=1 timeout 200 python tools/tc_phase_times.py bf16x3 > gpurun_out/r02e_phase.txt 2>&1; echo "rc=$?"; head -9 gpurun_out/r02e_phase.txt; grep -A6 "edge CTAs" gpurun_out/r02e_phase.txt | tail -6
