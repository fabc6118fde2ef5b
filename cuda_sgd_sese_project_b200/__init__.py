"""cuda_sgd_sese_project_b200 -- B200-native mini-batch SGD for linear regression, a drop-in for the
hot path of thvasilo/cuda-sgd-sese-project.  See DESIGN.md.  (The directory name uses underscores so
that it is importable; the reference's repository name has hyphens.)

Compute lives in libb200sgd.so (csrc/, hand-written sm_100a CUDA) behind include/b200sgd.h; this
package is the thin host-side mirror of the reference's interface used by tests, bench.py and the
torchrun launcher.  Importing it never triggers a CPU fallback: if the library is missing,
`_lib.load()` raises.
"""
from . import _lib
from ._lib import SgdError
from .engine import (SGDEngine, batch_geometry, device_count, get_filename_from_path, host_bucket,
                     host_shuffle, init_weights, read_csv, shard_range, write_csv, write_experiment_output)

__all__ = ["SGDEngine", "SgdError", "batch_geometry", "device_count", "get_filename_from_path",
           "host_bucket", "host_shuffle", "init_weights", "read_csv", "shard_range", "write_csv", "write_experiment_output", "_lib"]
