"""B200-native image->video sampling path (cINN inverse flow + 3-D conv decoder).

Drop-in for the reference's ``get_model.Model`` surface (get_model.py:10-103); the arithmetic runs in
hand-written sm_100a CUDA behind the C-ABI declared in ``include/i2v_b200.h``.
"""
__version__ = "0.1.0"
