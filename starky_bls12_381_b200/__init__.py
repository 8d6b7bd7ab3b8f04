"""starky_bls12_381_b200 -- B200-native (sm_100a) core of the starky prover used by
Electron-Labs/starky_bls12_381 (trace LDE -> Poseidon Merkle -> quotient -> FRI), behind a C ABI.

The product is the shared library ``libstarkyb200.so`` (CUDA kernels + host orchestration, see
``include/starky_b200.h``).  This Python package is only the harness side of that ABI (ctypes), used
by the tests and bench.py; it never computes anything itself and fails loudly when the library is missing.
"""
from .binding import (Params, ProofLayout, Proof, Context, StarkId, TraceLayout, Flags, lib, lib_path,  # noqa: F401
                      standard_params, SbError)
from .api import (StarkConfig, STARKS, prove, FP12MulStark, PairingPrecompStark, MillerLoopStark,  # noqa: F401
                  FinalExponentiateStark, ECCAggStark)
from .sharded import ShardPlan, shard_plan, commit_sharded, GpuBackend  # noqa: F401

__all__ = ["Params", "ProofLayout", "Proof", "Context", "StarkId", "TraceLayout", "Flags", "lib", "lib_path",
           "standard_params", "SbError", "StarkConfig", "STARKS", "prove", "FP12MulStark", "PairingPrecompStark", "MillerLoopStark",
           "FinalExponentiateStark", "ECCAggStark", "ShardPlan", "shard_plan", "commit_sharded",
           "GpuBackend"]
