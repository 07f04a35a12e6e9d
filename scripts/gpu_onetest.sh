set -x
mkdir -p gpurun_out
timeout 900 python -m pytest ${PYTEST_ARGS} -m gpu -x -q 2>&1 | tail -25
