"""ukbb_cardiac_b200: B200-native FCN segmentation deploy path of ukbb_cardiac.

Host-side Python mirrors the reference's deploy interface
(``common/deploy_network.py``); all arithmetic runs in ``libukbb_fcn.so``
(hand-written sm_100a CUDA behind the C ABI of ``include/ukbb_fcn.h``).
"""
__version__ = "0.1.0"
