"""Micro-benchmark of the tcgen05 conv kernel through the op-level test hook (timing = whole hook call
minus copies is not separable, so this uses the library profile of a full step instead).  Prefer:
    KKX_TC_DEBUG=<mask> KKX_TC_MULTI=<0|1> KKX_PROFILE_DETAIL=1 python tools/profile_step.py --batch 16
"""
