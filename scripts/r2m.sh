python -m pytest tests/test_gpu_device_domain.py tests/test_host_lbm.py tests/test_gpu_multi.py tests/test_zgpu_multi_next.py -m gpu -q --tb=short 2>&1 | tail -15
