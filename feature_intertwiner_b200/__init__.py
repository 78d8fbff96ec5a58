"""feature_intertwiner_b200 -- the Feature Intertwiner hot path (RoIAlign -> reliable / less-reliable split ->
class-mean buffer -> L2 / Sinkhorn intertwiner loss) as hand-written sm_100a CUDA behind a C ABI
(include/fi_b200.h), with the reference's own operator API on top.  See DESIGN.md.
"""
from ._lib import FiError, get_option, lib, library_path, set_option  # noqa: F401
from .roi_align import (CropAndResizeFunction, RoIAlign, crop_and_resize, crop_pair, crop_sets, crop_taps,  # noqa: F401
                        set_deterministic)
from .roi_pool import RoIPoolFunction, _RoIPooling  # noqa: F401
from .nms import nms, nms_batched, nms_presorted, proposal_decode, proposal_layer, pth_nms  # noqa: F401
from .ot import OptTrans, sinkhorn_loss  # noqa: F401
from .targets import detection_decode, detection_layer, mask_targets  # noqa: F401
from .intertwiner import (Dev, IntertwinerLoss, LevelSplit, assign_feat2cls, assign_feat2cls_multi, pyramid_roi_align, roi_level,  # noqa: F401
                          spatial_order, split_levels)
