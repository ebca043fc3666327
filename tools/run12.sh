cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/exp_guest.py > gpurun_out/r2_guest.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_guest_launches.csv python tools/exp_guest.py > /dev/null 2>&1
grep guest_energy gpurun_out/r2_guest_launches.csv | tail -3 >> gpurun_out/r2_guest.log
