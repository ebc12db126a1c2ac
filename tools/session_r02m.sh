python -m pytest tests/test_gpu_r02.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -12 | tee gpurun_out/r02m_pytest.log
