"""desman_b200 -- B200 (sm_100a) engine for DESMAN's haplotype-inference Gibbs sweep.

Drop-in surfaces (same names and semantics as the reference):
  desman_b200.sampletau            <- sampletau/sampletau.pyx
  desman_b200.HaploSNP_Sampler     <- desman/HaploSNP_Sampler.py
  desman_b200.Init_NMFT            <- desman/Init_NMFT.py
  bin/desman                       <- bin/desman
All arithmetic runs in hand-written CUDA kernels behind the C-ABI of include/desman_b200.h;
there is no CPU fallback.
"""
__version__ = "0.1.0"
