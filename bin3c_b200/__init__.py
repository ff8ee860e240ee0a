"""
bin3c_b200 -- B200-native implementation of the bin3C contact-map hot path
(pair accumulation -> filter mask -> site normalisation -> Knight-Ruiz balancing -> edge
weighting).  Importing the package does not need a GPU; running any stage does: the
hand-written sm_100a kernels in libbin3c_b200.so are the only implementation.
"""
__version__ = '0.1.0'
