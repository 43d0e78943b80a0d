"""csmri-refinement_b200: B200-native data-consistency (DC) hot path.

Only what the path needs: the CUDA kernels + C ABI (``csrc/``), the torch
custom ops (``ops``), and host-side mirrors of the reference interfaces that
sit on the path (``myfft.DataConsistencyInKspace``, ``recnet.RecNet``,
``undersampling``, ``parallel``).
"""
__version__ = '0.1.0'
