mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edges.py -q -m gpu -x --timeout=300 > gpurun_out/test_edges.log 2>&1; echo "edges exit $?"; tail -15 gpurun_out/test_edges.log
timeout 300 python tools/time_kernels.py edges_fused edge_map threshold > gpurun_out/time_edges.log 2>&1; tail -5 gpurun_out/time_edges.log
