"""sepfwi -- host-side mirror of the seisfwi/SEP-2023 interfaces for the elastic FWI hot path,
on top of libsepfwi.so (hand-written sm_100a CUDA behind a C ABI; no CPU fallback).

  sepfwi.engine       Propagator / ShotSpec: one persistent handle per GPU
  sepfwi.fwi_utils    paraGen / surveyGen / sourceGene / padding  (Ops/FWI/fwi_utils.py)
  sepfwi.fwi_ops      forward / backward / obscalc                (pybind module `fwi`, Src/Torch_Fwi.cpp)
  sepfwi.FWI_ops      FWIFunction, FWI, FWI_obscalc, FWI_Lame_Den, ... (Ops/FWI/FWI_ops.py)
  sepfwi.obj_wrapper  PyTorchObjective: scipy L-BFGS-B front-end (Ops/FWI/obj_wrapper.py)
  sepfwi.elasticSolver  GPU-backed elasticSolver              (DAS_Waveform_Modeling/src/elasticSolver.py)
"""
__version__ = "0.1.0"
