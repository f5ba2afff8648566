"""covid19uk_b200 -- B200-native MCMC likelihood hot path of chrism0dwk/covid19uk.

Python host code mirrors the reference's ``model_spec`` / ``inference`` API; the arithmetic runs in
hand-written sm_100a CUDA behind the C ABI of ``include/seir_b200.h`` (``libseir_b200.so``).
There is no CPU fallback.
"""
__version__ = "0.1.0"
