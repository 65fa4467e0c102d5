# round 2, remaining GPU tests on the final commit (those not covered by pass r2ak / r2am)
timeout 140 python -m pytest tests/test_plane_hooks.py tests/test_legacy.py tests/test_cli.py tests/test_reference_format.py tests/test_batch_driver.py -m gpu -x -q 2>&1 | tail -3
