"""trackingbench_slam_b200 — B200 (sm_100a) implementation of the trackingBench-SLAM feature-tracking
front end (pyramid -> FAST/ORB extract -> Hamming match) behind the C ABI of include/tb_capi.h.

The product is libtb_b200.so (hand-written CUDA, built by trackingbench_slam_b200.build).  This
package only binds it (capi) and mirrors the reference's host-side classes for Python callers
(frontend).  There is no CPU fallback: importing capi without the built library raises.
"""
from . import capi  # noqa: F401
from .capi import (Context, Frontend, DeviceVocabulary, Undistorter, init_undistort_rectify_map, MatcherParams, StereoParams, LKParams, RansacParams, Pinhole, TbError, KP_DTYPE, MATCH_DTYPE,  # noqa: F401
                   pyramid_geometry, lib, TB_OK, TB_E_INVALID, TB_E_CUDA, TB_E_CAPACITY, TB_E_UNSUPPORTED, TB_E_NOMEM)
from . import vocab  # noqa: F401
