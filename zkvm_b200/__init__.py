"""zkvm_b200 -- B200 (sm_100a) backend for ZkVM's one data-parallel hot path: the Ristretto255
variable-time multiscalar multiplication.  See DESIGN.md.  Importing the package does not load
CUDA; creating a `Context` does, and fails loudly without the compiled library or a GPU."""
from .ristretto import (CompressedRistretto, Context, InvalidPoint, PointTable, RistrettoPoint, Scalar, ZkError,
                        GROUP_ORDER, IDENTITY_BYTES, pick_window, batch_optional_multiscalar_mul,
                        batch_vartime_multiscalar_mul, MultiGpu, MultiGpuTable, host_register, host_unregister)

__all__ = ["CompressedRistretto", "Context", "InvalidPoint", "PointTable", "RistrettoPoint", "Scalar", "ZkError",
           "GROUP_ORDER", "IDENTITY_BYTES", "pick_window", "batch_optional_multiscalar_mul", "batch_vartime_multiscalar_mul",
           "MultiGpu", "MultiGpuTable", "host_register", "host_unregister"]
