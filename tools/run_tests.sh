cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/r2_tests.log 2>&1
