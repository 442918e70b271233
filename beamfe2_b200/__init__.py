"""beamfe2_b200 -- B200-native numeric core of BeamFE2 (2D Hermitian beam frames): element K/M,
banded assembly with support condensation, FP64 static solve and generalized modal eigenproblem as
hand-written sm_100a CUDA kernels behind a C ABI (libbfe2.so), batched over structure variants.

Drop-in modules mirror the reference package layout:
    from beamfe2_b200 import HermitianBeam_2D as HB, Structure, Node
Batched front end: beamfe2_b200.batch (solve_fused, HostPipeline), topology plan: beamfe2_b200.plan.
"""
__version__ = '0.1.0'
