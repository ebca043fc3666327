set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python tools/exp_build.py 0.10 2 > gpurun_out/p_build_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:grid_build_tiled -s 2 -c 1 -o gpurun_out/prof_r02_build python tools/exp_build.py 0.10 2 > gpurun_out/p_build_ncu.log 2>&1
python tools/exp_interp_c2.py 3 > gpurun_out/p_interp_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:interpolate_ -s 4 -c 2 -o gpurun_out/prof_r02_interp python tools/exp_interp_c2.py 3 > gpurun_out/p_interp_ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/p_bench_plain.json 2> gpurun_out/p_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c5 > gpurun_out/p_bench_ncu.json 2> gpurun_out/p_bench_ncu.err
ls -la gpurun_out/prof_r02_*.ncu-rep gpurun_out/r02_launches_bench.csv
