"""B200-native SSD300 detection-head hot path (default boxes, matching, MultiBox loss + gradient, decode, score,
NMS, TP/FP tallies) behind the call surface of rs1004/object-detection-torch2.

    from object_detection_torch2_b200.model import SSD                       # reference: from model import SSD
    from object_detection_torch2_b200.utils import calc_coordicate, ...      # reference: from utils import ...

Importing the package does not load CUDA; the first kernel call loads ``libssdhead.so`` and fails loudly when it
(or a CUDA device) is missing.
"""
__version__ = "0.1.0"
