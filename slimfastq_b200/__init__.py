"""slimfastq_b200: B200-native implementation of slimfastq's entropy-coding hot path.

`Codec` (api.py) is the host-side mirror of the reference's encode/decode loops over the C ABI in
include/sfq_b200.h; `container` reads the chunked .sfq container; `synth` generates the benchmark
inputs.  All coding runs in libsfq_b200.so's sm_100a kernels - there is no CPU path.
"""
from .api import (Codec, SfqError, decompressed_size, export_reference, import_reference,  # noqa: F401
                  is_reference_file, merge_containers, split_on_grid, split_records)

__all__ = ["Codec", "SfqError", "decompressed_size", "export_reference", "import_reference", "is_reference_file",
           "merge_containers", "split_on_grid", "split_records"]
