set -x
# ncu evidence of one round: launch list of a step, --set full captures of the three kernel groups.  P = file prefix.
P=${P:-r2c}
export GSR_TUNE=eigh_chunks=1
python tools/profile_step.py > gpurun_out/${P}_plain_step.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/${P}_launches_step.csv python tools/profile_step.py > gpurun_out/${P}_ncu_step.log 2>&1
python tools/tc_prof.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tcgemm3 -s 6 -c 3 -o gpurun_out/${P}_prof_tcgemm3 -f python tools/tc_prof.py > gpurun_out/${P}_ncu_tc.log 2>&1
# eigh_q_launch=2: Q formed by its own launch also with one launch per stage (ncu serialises kernels: no chunked pipeline)
export GSR_TUNE=eigh_chunks=1,eigh_q_launch=2
python tools/eigh_prof.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"tridiag_kernel|form_q|tql_" -s 5 -c 5 -o gpurun_out/${P}_prof_eigh -f python tools/eigh_prof.py > gpurun_out/${P}_ncu_eigh.log 2>&1
export GSR_TUNE=
GSR_NO_TIMES=1 python tools/ew_prof.py > /dev/null 2>&1 && GSR_NO_TIMES=1 ncu --set full --clock-control none --import-source on -k regex:"normalize_|sym_diag" -s 9 -c 9 -o gpurun_out/${P}_prof_ew -f python tools/ew_prof.py > gpurun_out/${P}_ncu_ew.log 2>&1
python tools/ew_prof.py > gpurun_out/${P}_ew_times.log 2>&1
ls -la gpurun_out/${P}_*
