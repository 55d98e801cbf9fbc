mkdir -p gpurun_out/${1:-tests}
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --durations=8 > gpurun_out/${1:-tests}/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${1:-tests}/pytest_gpu.log
tail -40 gpurun_out/${1:-tests}/pytest_gpu.log
