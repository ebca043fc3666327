set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
TAG=${1:-b8}
python tools/exp_build.py 0.10 2 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:grid_build_tiled -s 2 -c 1 -o gpurun_out/prof_${TAG} python tools/exp_build.py 0.10 2 > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
