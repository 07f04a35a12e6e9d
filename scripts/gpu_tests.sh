set -x
mkdir -p gpurun_out
TAG=${TAG:-t}
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ${PYTEST_ARGS} 2>&1 | tail -30 > gpurun_out/${TAG}_pytest.log
cat gpurun_out/${TAG}_pytest.log
