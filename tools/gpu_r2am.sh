# round 2, last check after the framing checks moved into sfq_container.h: smoke(), the error-path test, two parity tests
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "errors or illumina_chunks or reference_samples" 2>&1 | tail -2
