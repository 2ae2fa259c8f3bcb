"""vlgae_b200 -- B200-native (sm_100a) implementation of VLGAE's structured-inference hot path.

Only what the path needs lives here:

* ``csrc/``           hand-written CUDA kernels + the C ABI (``include/vlgae_b200.h``)
* ``_lib.py``         ctypes binding of ``libvlgae_b200.so`` (fails loudly when it is missing)
* ``ops.py``          thin tensor-level wrappers (device pointers + current CUDA stream)
* ``torch_struct/``   mirror of the reference's ``src/model/torch_struct`` operator API
                      (``DMV1o``, ``DependencyCRF``, constants, ``semirings.semirings.NEGINF``)
* ``alignment.py``    the ``gather_logit`` implementation group of ``src/model/joint.py``
* ``sharding.py``     sentence sharding of a batch across the GPUs of one box

There is no CPU fallback: every operator raises if the CUDA library or a CUDA device is absent.
"""
__version__ = "0.1.0"
