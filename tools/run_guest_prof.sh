cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/exp_guest.py > gpurun_out/guest_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:guest_energy -s 1 -c 1 -o gpurun_out/prof_r02_guest3 python tools/exp_guest.py > gpurun_out/guest_ncu.log 2>&1
cat gpurun_out/guest_plain.log
