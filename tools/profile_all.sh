set -x
export GSR_TUNE=eigh_chunks=1
python tools/profile_step.py > gpurun_out/r2b_plain_step.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2b_launches_step.csv python tools/profile_step.py > gpurun_out/r2b_ncu_step.log 2>&1
python tools/tc_prof.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tcgemm3 -s 6 -c 3 -o gpurun_out/r2b_prof_tcgemm3 -f python tools/tc_prof.py > gpurun_out/r2b_ncu_tc.log 2>&1
python tools/eigh_prof.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tridiag_kernel|tql_" -s 4 -c 4 -o gpurun_out/r2b_prof_eigh -f python tools/eigh_prof.py > gpurun_out/r2b_ncu_eigh.log 2>&1
python tools/ew_prof.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"normalize_|sym_diag" -s 8 -c 8 -o gpurun_out/r2b_prof_ew -f python tools/ew_prof.py > gpurun_out/r2b_ncu_ew.log 2>&1
ls -la gpurun_out/r2b_*
